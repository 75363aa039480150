#!/usr/bin/env python
"""bench.py -- 1080p motion-mask frames/s of the B200 hot path (BASELINE.json metric), one JSON line on rank 0.

A "step" is one pass of the hot path (pyramid -> LK -> egomotion fit -> fused warp/diff/threshold/morphology) over one
batch of `--batch` frame pairs of the 1920x1080 synthetic sequence with global affine camera motion plus moving blobs
(BASELINE.json configs[1], SURVEY.md 8d "C2").  `value` = pairs/s with the frames already resident in HBM; `e2e` = the
same metric through the C ABI with HOST (pinned) buffers, H2D of the frames and D2H of masks/flow/H inside the timed
region.  N > 1: one process per GPU (torchrun), independent camera streams per rank (every rank its own copy of the same synthetic sequence: LK work is data dependent, so
only identical content keeps the per-GPU work of a weak-scaling run fixed), no collective
on the frame path; NCCL only for the barrier / max-over-ranks / stats gather.

  python bench.py --gpus 1 --steps 20 --warmup 3
  python bench.py --impl reference ...    # the reference's CPU path (OpenCV via cv2, all host threads) on rank 0
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "1080p motion-mask frames/sec"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="frame pairs per step (round 2, value / e2e frames/s: 32: 5 866 / 5 770, 48: 5 928 / 5 868, "
                                                          "64: 5 964 / 5 908; the head and tail of a batch are paid once per step)")
    ap.add_argument("--pixel-step", type=int, default=10, help="grid step (launch-file default 10; 1 = dense)")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--ref-pairs", type=int, default=4, help="pairs per step of the reference arm (bounded sample)")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="pairs", choices=["pairs", "live"],
                    help="pairs = the motion-mask chain (BASELINE metric); live = the node's imageCallback path "
                         "(trajectory window -> fitSubspace -> clusterEuclidean -> boxes), a secondary line")
    ap.add_argument("--num-motions", type=int, default=2)
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the short secondary measurements (dense 1080p, VGA, 4K LK, 4K VarFlow chain, VarFlow 1080p, live path) "
                         "the default N = 1 run appends under \"secondary\"")
    ap.add_argument("--lean", action="store_true", help="skip the extra figures (two streams per GPU, 1-thread CPU) -- used by the secondary runs")
    ap.add_argument("--no-pin", action="store_true", help="do not pin the rank to the host cores next to its GPU")
    ap.add_argument("--flow-engine", default="lk", choices=["lk", "varflow"],
                    help="flow feeding the egomotion fit: grid LK (default) or the dense variational flow VarFlow::CalcFlow "
                         "sampled at the grid (BASELINE configs[2]: 4-level pyramid, dense flow + homography)")
    return ap.parse_args()


def workload_name(a):
    return "C2 %dx%d synthetic affine camera + 3 moving blobs, pixel_step=%d, min_vector_size=0.2, RANSAC homography%s" % (
        a.width, a.height, a.pixel_step, ", dense VarFlow engine (max_level 4)" if a.flow_engine == "varflow" else "")


def copies_resident(a):
    """inputs larger than L2: the timed loop rotates over R resident copies of the batch (R * (B+1) frames > 300 MB)"""
    return max(2, int(np.ceil(300e6 / ((a.batch + 1) * a.width * a.height))))


def config_dict(a, world):
    """`config` of the JSON line -- the SAME dict for the B200 arm and the reference arm (the reference arm times a bounded
    sample of this workload; what the sample was is said in its cpu_baseline.sample)."""
    gx = (a.width + a.pixel_step - 1) // a.pixel_step
    gy = (a.height + a.pixel_step - 1) // a.pixel_step
    R = copies_resident(a)
    return {"workload": workload_name(a), "pairs_per_step": a.batch, "grid_points": gx * gy, "streams": world,
            "l2": "inputs rotate over %d resident copies (%.0f MB > 126 MB L2)" % (R, R * (a.batch + 1) * a.width * a.height / 1e6),
            "parallelism": "independent camera streams, one per GPU, no frame-path collective"}


def pin_to_gpu_cores(local, world):
    """N > 1: pin this rank to host cores next to its GPU (the cores of the GPU's NUMA node that the container may use, else
    the container's cores), a disjoint slice per rank: a step is ~100 API calls per 7 ms, and launching threads that migrate
    or share cores show up as launch jitter at N >= 2.  Returns what was done (reported in the JSON line)."""
    try:
        import torch
        avail = sorted(os.sched_getaffinity(0))

        def node_cores(dev):
            pr = torch.cuda.get_device_properties(dev)
            path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            cpus = []
            for part in open(path).read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus += list(range(int(lo), int(hi or lo) + 1))
            near = sorted(set(cpus) & set(avail))
            return tuple(near) if near else tuple(avail)

        bases = [node_cores(d) for d in range(world)]
        mine = bases[local]
        sharers = [d for d in range(world) if bases[d] == mine]
        per = max(1, len(mine) // len(sharers))
        k = sharers.index(local)
        cores = list(mine[k * per:(k + 1) * per]) or list(mine)
        os.sched_setaffinity(0, cores)
        return {"cores": len(cores), "first": cores[0], "last": cores[-1], "numa_local": set(cores) <= set(node_cores(local)) and len(mine) < len(avail)}
    except Exception as e:          # no sysfs entry, no permission: run unpinned
        return {"cores": None, "error": str(e)[:80]}


def make_frames(a, rank, n):
    """Every rank gets its own copy of the SAME sequence: the LK iteration count depends on the image content (r02: the seed
    1234 + 1 sequence needs 4.6 % more GPU time than seed 1234 on the same GPU, which read as a flat scaling loss at N >= 2)."""
    from motion_detection_b200 import synth
    frames, _ = synth.sequence(a.width, a.height, n, seed=1234)
    return frames


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        try:                         # the physical GPU, whatever CUDA_VISIBLE_DEVICES maps `index` to
            import torch
            self.index = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
        except Exception:
            pass
        self.rows = []
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        if os.environ.get("MD_BENCH_NO_SAMPLER"):        # diagnostic runs only: does the polling itself disturb the GPU?
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", os.environ.get("MD_BENCH_SAMPLER_MS", "20")],
                                         stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference(a, rank, world):
    """The reference's own CPU implementation of the path (OpenCV routines through cv2, all host threads)."""
    if rank != 0:
        return
    from oracle import cv_chain
    frames = make_frames(a, 0, a.ref_pairs + 1)

    def step():
        for b in range(a.ref_pairs):
            cv_chain.process_pair(frames[b], frames[b + 1], pixel_step=a.pixel_step, min_vector_size=0.2, seed=1 + b)

    for _ in range(a.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    val = a.ref_pairs * a.steps / dt
    sample = "%d steps x %d pairs (a bounded sample of the %d-pair step) of the same %s" % (a.steps, a.ref_pairs, a.batch, workload_name(a))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/f32/f64 (integer samples, f32 sums, f64 homography)", "data": "synthetic",
        "config": config_dict(a, max(world, 1)),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cv_chain.threads(), "kind": "port", "sample": sample,
                         "host": "cv2 %s" % ("present" if cv_chain.have_cv2() else "absent -> plain-C oracle")},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mpx_per_s": val * a.width * a.height / 1e6,
    }
    print(json.dumps(line), flush=True)


def run_live(a, rank, local, world):
    """Secondary workload: the node's live path.  A step = `--batch` imageCallbacks, each = one NEW frame pushed into the
    ring (gray conversion + pyramid once) + md_window_detect over the newest F = 2 * num_motions + 1 frames."""
    import torch
    import torch.distributed as dist
    from motion_detection_b200 import capi, streams
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    w, h, B, nm = a.width, a.height, a.batch, a.num_motions
    F = 2 * nm + 1
    name = "live path (imageCallback): %dx%d C2 sequence, pixel_step=%d, window F=%d, sigma=0.5, distance_threshold=50" % (
        w, h, a.pixel_step, F)
    ctx = capi.Context(width=w, height=h, max_batch=F - 1, pixel_step=a.pixel_step, min_vector_size=0.2, seed=1, device=local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    clip = make_frames(a, rank, B + 1)
    # ping-pong order keeps every pushed frame a real neighbour of the previous one
    order = list(range(B + 1)) + list(range(B - 1, 0, -1))
    R = max(2, int(np.ceil(300e6 / ((B + 1) * w * h))))
    sets = torch.empty((R, B + 1, h, w), dtype=torch.uint8, device=dev)
    for r in range(R):
        sets[r].copy_(torch.from_numpy(clip))
    lp = capi.MdLiveParams()
    capi.lib().md_live_params_default(C.byref(lp))
    lp.num_motions = nm
    res = capi.MdLiveResult()
    state = {"k": 0, "cb": 0, "traj": 0, "boxes": 0}

    def callback(push):
        k = state["k"]
        state["k"] += 1
        fill = push(order[k % len(order)], k)
        if fill >= F:
            lp.seed = 1 + state["cb"]
            ctx.raw_window_detect(lp, res, capi.MD_MEM_DEVICE)
            state["cb"] += 1
            state["traj"] += res.num_trajectories
            state["boxes"] += res.num_clusters

    def push_dev(f, k):
        return ctx.raw_window_push(sets[(k // len(order)) % R, f].data_ptr(), 1, w, capi.MD_MEM_DEVICE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(push, steps):
        for _ in range(steps * B):
            callback(push)

    run(push_dev, max(a.warmup, 1))
    barrier()
    l0 = ctx.stats()["kernel_launches"]
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.15)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    cb0 = state["cb"]
    e0.record(stream)
    run(push_dev, a.steps)
    e1.record(stream)
    barrier()
    ms = streams.max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    launches = ctx.stats()["kernel_launches"] - l0
    ncb = state["cb"] - cb0
    value = world * ncb / (ms * 1e-3)
    assert state["traj"] > 0

    e2e = None
    if not a.no_e2e:
        rgb = torch.from_numpy(np.repeat(clip[..., None], 3, axis=3)).pin_memory()       # the node hands over rgb8 (node.cpp:271)
        P = ctx.P
        pin = dict(boxes=torch.empty((P, 4), dtype=torch.int32).pin_memory(), sizes=torch.empty((P,), dtype=torch.int32).pin_memory(),
                   opts=torch.empty((P, 2), dtype=torch.float32).pin_memory(), labels=torch.empty((P,), dtype=torch.int32).pin_memory())
        hres = capi.MdLiveResult()
        hres.boxes, hres.cluster_sizes = pin["boxes"].data_ptr(), pin["sizes"].data_ptr()
        hres.outlier_points, hres.labels = pin["opts"].data_ptr(), pin["labels"].data_ptr()
        ctx.window_reset()
        st2 = {"k": 0, "cb": 0, "d2h": 0}

        def hcallback():
            k = st2["k"]
            st2["k"] += 1
            fill = ctx.raw_window_push(rgb[order[k % len(order)]].data_ptr(), 3, 3 * w, capi.MD_MEM_HOST)
            if fill >= F:
                lp.seed = 1 + st2["cb"]
                ctx.raw_window_detect(lp, hres, capi.MD_MEM_HOST)
                st2["cb"] += 1
                st2["d2h"] += 20 + hres.num_outliers * 12 + hres.num_clusters * 20

        for _ in range(max(a.warmup, 1) * B):
            hcallback()
        barrier()
        c0, d0 = st2["cb"], st2["d2h"]
        t0 = time.perf_counter()
        for _ in range(a.steps * B):
            hcallback()
        torch.cuda.synchronize()
        dt = streams.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * (st2["cb"] - c0) / dt, "unit": "callbacks/s", "h2d_bytes_per_step": B * 3 * w * h,
               "d2h_bytes_per_step": int((st2["d2h"] - d0) / max(a.steps, 1))}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle import cv_chain
        n = 0
        cv_chain.live_callback(clip[:F], pixel_step=a.pixel_step, num_motions=nm)
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < a.cpu_baseline_seconds and n < 500:
            i = n % (B + 2 - F)
            cv_chain.live_callback(clip[i:i + F], pixel_step=a.pixel_step, num_motions=nm, seed=1 + n)
            n += 1
        dt = time.perf_counter() - t0
        cpu = {"value": n / dt, "unit": "callbacks/s", "cores": cv_chain.threads(), "kind": "port",
               "sample": "%d callbacks of the same workload in %.1f s (cv2 OpenCV routines re-run over the whole window per "
                         "callback like the reference, oracle fitSubspace + clusterEuclidean)" % (n, dt)}
    if rank == 0:
        N, P = w * h, ctx.P
        alg = 1.3333 * N + (F - 1) * (2.6667 * N + 9 * P) + 8.0 * P * F      # new pyramid + F-1 LK passes + trajectories
        peak, peak_src = peaks()
        gbs = alg * ncb / (ms * 1e-3) / 1e9
        line = {
            "metric": "1080p live-path callbacks/sec (secondary workload)", "value": value, "unit": "callbacks/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/f32/f64", "data": "synthetic",
            "config": {"workload": name, "callbacks_per_step": B, "grid_points": P, "streams": world,
                       "l2": "frames rotate over %d resident copies of the clip (%.0f MB > 126 MB L2)" % (R, R * (B + 1) * N / 1e6)},
            "roofline": {"bound": "hbm", "kernel": "whole callback (k_lk_phase + 3 x k_lk_tma dominate)", "achieved": gbs, "peak": peak,
                         "unit": "GB/s", "frac": gbs / peak, "traffic": None, "peak_source": peak_src,
                         "note": "LK is instruction-issue bound (see DESIGN.md); algorithmic bytes per callback = %.1f MB" % (alg / 1e6)},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "live_stats": {"callbacks": ncb, "mean_trajectories": state["traj"] / max(state["cb"], 1),
                           "mean_boxes": state["boxes"] / max(state["cb"], 1)},
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


SECONDARY = [
    # (key, BASELINE config it stands for, extra flags)  -- a few steps each, every one a full bench line of its own
    ("dense_1080p", "configs[1] dense: every pixel tracked (pixel_step 1, 2 073 600 points per pair)",
     ["--pixel-step", "1", "--batch", "2", "--steps", "3", "--warmup", "3", "--cpu-baseline-seconds", "4"]),
    ("c1_vga", "configs[0]: 640x480", ["--width", "640", "--height", "480", "--batch", "128", "--steps", "10", "--warmup", "3",
                                       "--cpu-baseline-seconds", "4"]),
    ("c3_4k_lk", "configs[2] with the LK engine: 3840x2160, 6 pyramid levels", ["--width", "3840", "--height", "2160", "--batch", "8",
                                                                                "--steps", "6", "--warmup", "3", "--cpu-baseline-seconds", "4"]),
    ("c3_4k_varflow", "configs[2] as written: 3840x2160, dense variational flow (VarFlow, max_level 4) + homography egomotion",
     ["--width", "3840", "--height", "2160", "--flow-engine", "varflow", "--batch", "8", "--steps", "3", "--warmup", "3",
      "--cpu-baseline-seconds", "1"]),
    ("varflow_1080p", "configs[1] with the dense variational flow engine", ["--flow-engine", "varflow", "--batch", "8", "--steps", "3",
                                                                            "--warmup", "3", "--cpu-baseline-seconds", "1"]),
    ("live_1080p", "configs[4]-style live path: imageCallback (window of 5 frames -> trajectories -> fitSubspace -> clusters)",
     ["--workload", "live", "--batch", "8", "--steps", "4", "--warmup", "3", "--cpu-baseline-seconds", "4"]),
]


def run_secondary(a):
    """The other BASELINE configurations, a few steps each, every one measured by this same script in a child process (own
    contexts, own timing rules: >= 3 warm-up steps, inputs larger than L2, CUDA events, e2e with host buffers, CPU beside it)."""
    out = {}
    for key, what, flags in SECONDARY:
        cmd = [sys.executable, os.path.abspath(__file__), "--gpus", "1", "--no-secondary", "--lean"] + flags
        t0 = time.perf_counter()
        try:
            pr = subprocess.run(cmd, capture_output=True, text=True, timeout=420)
            rows = [l for l in pr.stdout.splitlines() if l.startswith("{")]
            if pr.returncode != 0 or not rows:
                out[key] = {"what": what, "error": (pr.stderr or pr.stdout)[-300:]}
                continue
            d = json.loads(rows[-1])
            keep = {k: d.get(k) for k in ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "config", "roofline", "stages",
                                          "cpu_baseline", "e2e", "gpu_launches", "lk_work", "mpx_per_s", "live_stats")}
            keep["what"] = what
            keep["flags"] = " ".join(flags)
            keep["wall_s"] = round(time.perf_counter() - t0, 1)
            out[key] = keep
        except Exception as e:                      # a secondary line never takes the headline down with it
            out[key] = {"what": what, "error": str(e)[-300:]}
    return out


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference" and a.workload == "pairs":
        run_reference(a, rank, world)
        return
    if a.workload == "live":
        if a.impl == "reference":
            raise SystemExit("--impl reference times the BASELINE workload (pairs); the live CPU figure is the cpu_baseline of --workload live")
        run_live(a, rank, local, world)
        return

    import torch
    import torch.distributed as dist
    from motion_detection_b200 import capi, streams

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    pinned = pin_to_gpu_cores(local, world) if world > 1 and not a.no_pin else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    w, h, B = a.width, a.height, a.batch

    engine = capi.MD_FLOW_VARFLOW if a.flow_engine == "varflow" else capi.MD_FLOW_LK
    ctx = capi.Context(width=w, height=h, max_batch=B, pixel_step=a.pixel_step, min_vector_size=0.2, seed=1, device=local,
                       flow_engine=engine)
    P = ctx.P
    # all kernels AND the timing events go on one explicit (non-default) torch stream: handle 0 would mean
    # "the context's own stream" to md_set_stream and the events would not bracket the work
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)

    frames_np = make_frames(a, rank, B + 1)
    frame_bytes = w * h
    # inputs larger than L2: rotate over R resident copies of the batch (R * (B+1) * 2 MB > 126 MB)
    R = copies_resident(a)
    host = torch.from_numpy(frames_np)
    sets = torch.empty((R, B + 1, h, w), dtype=torch.uint8, device=dev)
    for r in range(R):
        sets[r].copy_(host)
    d_mask = torch.empty((B, h, w), dtype=torch.uint8, device=dev)
    d_next = torch.empty((B, P, 2), dtype=torch.float32, device=dev)
    d_status = torch.empty((B, P), dtype=torch.uint8, device=dev)
    d_keep = torch.empty((B, P), dtype=torch.uint8, device=dev)
    d_H = torch.empty((B, 9), dtype=torch.float64, device=dev)
    d_nv = torch.empty((B,), dtype=torch.int32, device=dev)
    d_inl = torch.empty((B,), dtype=torch.int32, device=dev)
    outs = capi.MdOutputs(d_next.data_ptr(), d_status.data_ptr(), d_keep.data_ptr(), d_H.data_ptr(), d_nv.data_ptr(),
                          d_inl.data_ptr(), d_mask.data_ptr(), w, w * h)

    def step(i):
        ctx.raw_process_batch(sets[i % R].data_ptr(), 1, w, frame_bytes, B + 1, False, outs, capi.MD_MEM_DEVICE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # One-time setup, like the allocations above: md_process_batch runs a buffer set eagerly the first time it sees it, captures its
    # CUDA graph the second time and replays it from the third.  Every one of the R resident copies is taken through that before the
    # W warm-up steps, so that the timed region holds steady-state replays only (otherwise captures land inside it for W < 2 R).
    priming = 2 * R if a.flow_engine != "varflow" else 0             # the variational engine is not graph captured
    for i in range(priming):
        step(i)
    for i in range(a.warmup):
        step(i)
    barrier()
    st0 = ctx.stats()
    l0 = st0["kernel_launches"]
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.15)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(a.steps):
        step(a.warmup + i)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    st1 = ctx.stats()
    launches = st1["kernel_launches"] - l0
    lk_work = (st1["lk_iterations"] - st0["lk_iterations"], st1["lk_levels"] - st0["lk_levels"])
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    ms_ranks = [ms]
    if world > 1:
        tl = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(tl, t)
        ms_ranks = [float(x.item()) for x in tl]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * a.steps / (ms_max * 1e-3)

    # sanity of the timed work: every pair produced an egomotion fit and a mask
    inl = d_inl.cpu().numpy()
    assert (inl > 0).all() or a.flow_engine == "varflow", "egomotion fit failed inside the timed region"

    # ---- per-stage CUDA-event timing (same workload, events on the launching stream) -> roofline of the dominant kernel
    ctx.profile(True)
    stage_ms = np.zeros(4)
    nprof = max(3, min(10, a.steps))
    for i in range(nprof):
        step(i)
        stage_ms += np.array(ctx.profile_read())
    ctx.profile(False)
    stage_ms /= nprof
    N = w * h
    vf = a.flow_engine == "varflow"
    # algorithmic bytes per stage.  VarFlow: the two frames in, U and V out = 10 N per pair; the multigrid
    # solver's own streaming model (every smoothing sweep reading its operands once: ~800 N) is reported beside it.
    stage_bytes = [1.3333 * N * (B + 1), (10.0 * N if vf else 2.6667 * N + 9 * P) * B, 9.0 * P * B, 3.0 * N * B]
    names = ["K1 pyramid+scharr (k_level0/k_pyrdown/k_scharr)",
             "K2 dense variational flow VarFlow::CalcFlow (k_vf_*: presmooth, derivatives, Gauss-Seidel wavefront V-cycles)" if vf else
             "K2 pyramidal LK on the tracked grid (k_window_sums_ring = phase planes + window sums, k_lk_phase)",
             "K3 egomotion (k_keep_count/k_scan/k_compact/k_hypotheses/k_score/k_accum/k_solve)",
             "K4 fused warp+diff+threshold+erode+dilate (k_mask, TMA-staged tiles)"]
    peak, peak_src = peaks()
    stages = []
    for n_, m_, by in zip(names, stage_ms, stage_bytes):
        gbs = by / (m_ * 1e-3) / 1e9 if m_ > 0 else 0.0
        stages.append({"kernel": n_, "ms": float(m_), "algorithmic_bytes": by, "gbs": gbs, "frac": gbs / peak})
    dom = int(np.argmax(stage_ms))
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            per_pair = json.load(open(tp)).get(["k_pyramid", "k_lk", "k_ego", "k_mask"][dom])
            traffic = per_pair * B if per_pair is not None else None     # ncu dram bytes per pair x pairs per launch
        except Exception:
            traffic = None
    # K2's work in window-tap evaluations (1600 taps per iteration, counted on the device in the timed region)
    lk_taps = None
    if lk_work[0] > 0:
        per_point = lk_work[0] / (P * B * a.steps)
        lk_taps = {"iterations_per_point": per_point, "levels_iterated_per_point": lk_work[1] / (P * B * a.steps),
                   "tap_evaluations_per_pair": 1600.0 * lk_work[0] / (B * a.steps),
                   "tap_evaluations_per_s_device_resident": 1600.0 * lk_work[0] * world / (ms_max * 1e-3)}
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": stages[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": stages[dom]["frac"], "traffic": traffic, "peak_source": peak_src, "peak_spec": 8000.0,
                "frac_of_spec": stages[dom]["gbs"] / 8000.0,
                "share_of_step": float(stage_ms[dom] / stage_ms.sum()),
                "streaming_model": ({"bytes_per_pair": 800.0 * N, "frac": 800.0 * N * B / (stage_ms[1] * 1e-3) / 1e9 / peak,
                                     "note": "VarFlow multigrid: ~800 N bytes if every smoothing sweep streamed its operands from HBM once"}
                                    if vf else None),
                "note": "LK with a 40x40 window is ALU/shared-memory bound by construction (1600 taps x levels x iterations "
                        "per point); its HBM fraction is small by design, see stages[] for the HBM-bound kernels (K1, K4)"}

    # LK is bound by instruction issue, not by bytes: report that roofline beside the HBM one.  Warp instructions per tracked point =
    # static instruction counts per SASS region of k_lk_phase (profiles/traffic.json "k_lk_instr_model", from the committed ncu
    # source-page capture) x the iterations per point COUNTED ON THE DEVICE in this run's timed region; the time is measured here.
    try:
        tj = json.load(open(tp))
        model, ipp_ncu = tj.get("k_lk_instr_model"), tj.get("k_lk_warp_instr_per_point")
    except Exception:
        model, ipp_ncu = None, None
    if model and lk_taps and a.pixel_step == 10 and (w, h) == (1920, 1080) and not vf:
        ipp = (model["per_point"] + model["per_level_iterated"] * lk_taps["levels_iterated_per_point"]
               + model["per_iteration"] * lk_taps["iterations_per_point"])
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        peak_issue = 148 * 4 * sm_mhz * 1e6                       # one warp instruction per scheduler per clock
        ach = ipp * P * B / (stage_ms[1] * 1e-3)
        roofline["issue"] = {"bound": "instruction issue (148 SMs x 4 schedulers x SM clock)", "achieved": ach / 1e9, "peak": peak_issue / 1e9,
                             "unit": "G warp-instr/s", "frac": ach / peak_issue, "warp_instr_per_point": ipp,
                             "warp_instr_per_point_ncu": ipp_ncu,
                             "thread_instr_per_window_tap": 32.0 * ipp / (1600.0 * lk_taps["iterations_per_point"]),
                             "source": "k_lk_phase only: per-region static SASS counts (profiles/r02_lk_phase_regions.txt) x levels and iterations per "
                                       "point counted on the device in this run; the time is the whole K2 stage (planes + window sums + LK), "
                                       "so frac is a lower bound of the LK kernel's own issue utilisation (79 % under ncu)"}

    # ---- e2e: C ABI with HOST (pinned) buffers, H2D + D2H inside the timed region
    def measure_e2e(packed):
        ctx2 = capi.Context(width=w, height=h, max_batch=B, pixel_step=a.pixel_step, min_vector_size=0.2, seed=1, device=local,
                            flow_engine=engine, mask_packed=1 if packed else 0)
        mrow = (w + 7) // 8 if packed else w
        pin_frames = torch.empty((B + 1, h, w), dtype=torch.uint8).pin_memory()
        pin_frames.copy_(host)
        pin = dict(mask=torch.empty((B, h, mrow), dtype=torch.uint8).pin_memory(),
                   nxt=torch.empty((B, P, 2), dtype=torch.float32).pin_memory(),
                   st=torch.empty((B, P), dtype=torch.uint8).pin_memory(),
                   keep=torch.empty((B, P), dtype=torch.uint8).pin_memory(),
                   H=torch.empty((B, 9), dtype=torch.float64).pin_memory(),
                   nv=torch.empty((B,), dtype=torch.int32).pin_memory(),
                   inl=torch.empty((B,), dtype=torch.int32).pin_memory())
        houts = capi.MdOutputs(pin["nxt"].data_ptr(), pin["st"].data_ptr(), pin["keep"].data_ptr(), pin["H"].data_ptr(),
                               pin["nv"].data_ptr(), pin["inl"].data_ptr(), pin["mask"].data_ptr(), mrow, mrow * h)
        # Chained stream: the context keeps the pyramid of the last frame, every step pushes B NEW frames.  The clip
        # is played forward (f1..fB) and backward (fB-1..f0) alternately so that every pair is a real consecutive pair.
        pin_fwd = pin_frames[1:]
        pin_bwd = torch.empty((B, h, w), dtype=torch.uint8).pin_memory()
        pin_bwd.copy_(torch.flip(host[:B], dims=[0]))
        ctx2.raw_process_batch(pin_frames.data_ptr(), 1, w, frame_bytes, 2, False, houts, capi.MD_MEM_HOST)   # primes f0 -> f1
        ctx2.raw_process_batch(pin_frames.data_ptr(), 1, w, frame_bytes, 1, True, houts, capi.MD_MEM_HOST)    # back at f0
        nstep = [0]

        def hstep():
            src = pin_fwd if nstep[0] % 2 == 0 else pin_bwd
            nstep[0] += 1
            ctx2.raw_process_batch(src.data_ptr(), 1, w, frame_bytes, B, True, houts, capi.MD_MEM_HOST)

        for _ in range(max(a.warmup, 3) + (max(a.warmup, 3) % 2)):      # an even number of warm-up steps: the timed loop starts forward
            hstep()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            hstep()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        h2d = B * frame_bytes
        d2h = B * (mrow * h + P * 8 + P + P + 72 + 4 + 4)
        res = {"value": world * B * a.steps / float(t.item()), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "graph_replays": ctx2.stats()["graph_replays"]}
        assert int(pin["inl"].min()) > 0
        ctx2.close()
        return res

    # ---- e2e through streams.BatchPipeline: the same sequence, batches alternating over two contexts of this GPU as asynchronous
    # host-buffer calls (MD_MEM_HOST_ASYNC; unchained, one overlap frame per batch -> B + 1 uploads per step), results identical to
    # the chained single context (tests/test_gpu_stream.py).  Every step's H2D and D2H are inside the timed region; the host waits
    # for (and reads) batch k - 1 after submitting batch k.
    def measure_e2e_pipeline(lanes=2):
        pipe = streams.BatchPipeline(lanes=lanes, width=w, height=h, max_batch=B, pixel_step=a.pixel_step, min_vector_size=0.2, seed=1,
                                     device=local, flow_engine=engine)
        clips = [torch.empty((B + 1, h, w), dtype=torch.uint8).pin_memory() for _ in range(2)]
        clips[0].copy_(host)                                     # f0 .. fB
        clips[1].copy_(torch.flip(host, dims=[0]))               # fB .. f0: played alternately, every pair is a real consecutive pair
        pins, houts = [], []
        for _ in range(lanes):
            pin = dict(mask=torch.empty((B, h, w), dtype=torch.uint8).pin_memory(),
                       nxt=torch.empty((B, P, 2), dtype=torch.float32).pin_memory(),
                       st=torch.empty((B, P), dtype=torch.uint8).pin_memory(),
                       keep=torch.empty((B, P), dtype=torch.uint8).pin_memory(),
                       H=torch.empty((B, 9), dtype=torch.float64).pin_memory(),
                       nv=torch.empty((B,), dtype=torch.int32).pin_memory(),
                       inl=torch.empty((B,), dtype=torch.int32).pin_memory())
            pins.append(pin)
            houts.append(capi.MdOutputs(pin["nxt"].data_ptr(), pin["st"].data_ptr(), pin["keep"].data_ptr(), pin["H"].data_ptr(),
                                        pin["nv"].data_ptr(), pin["inl"].data_ptr(), pin["mask"].data_ptr(), w, w * h))
        low = [1 << 30]

        def run(nsteps):
            last = None
            for i in range(nsteps):
                t = pipe.submit(clips[pipe.batches % 2].data_ptr(), 1, w, frame_bytes, B + 1, houts[pipe.batches % lanes])
                if last is not None:
                    pipe.wait(last)
                    low[0] = min(low[0], int(pins[pipe.lane_of(last)]["inl"].min()))      # the host reads the finished batch
                last = t
            pipe.wait(last)
            low[0] = min(low[0], int(pins[pipe.lane_of(last)]["inl"].min()))

        run(2 * lanes * 2 + 2 * lanes)                            # every lane sees both clips eager, captured and replayed
        torch.cuda.synchronize()
        barrier()
        nst = a.steps + (a.steps % 2)
        t0 = time.perf_counter()
        run(nst)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert low[0] > 0
        res = {"value": world * B * nst / float(t.item()), "unit": UNIT, "h2d_bytes_per_step": (B + 1) * frame_bytes,
               "d2h_bytes_per_step": B * (w * h + P * 8 + P + P + 72 + 4 + 4), "steps": nst,
               "api": "streams.BatchPipeline, %d contexts fed alternately with MD_MEM_HOST_ASYNC batches of one sequence" % lanes,
               "graph_replays": pipe.stats()["graph_replays"]}
        pipe.close()
        return res

    e2e = None
    e2e_packed = None
    e2e_one = None
    if not a.no_e2e:
        e2e = measure_e2e(False)
        e2e["api"] = "md_process_batch(MD_MEM_HOST), one context, chained"
        if not vf:
            e2e_one = e2e
            e2e_pipe = measure_e2e_pipeline(2)
            if e2e_pipe["value"] > e2e_one["value"]:
                e2e = e2e_pipe
        if not a.lean and not vf:
            # the same with md_config.mask_packed (1 bit per mask pixel on the way back: the u8 mask is 91 % of the D2H bytes)
            e2e_packed = measure_e2e(True)

    # ---- host link with all ranks copying at once: what bounds e2e when N ranks share one host (64 MB pinned copies each way, both
    # directions in flight together like in the pipelined host path); the slowest rank is reported
    host_link = None
    if not a.no_e2e:
        nb = 64 << 20
        hp_in, hp_out = torch.empty(nb, dtype=torch.uint8).pin_memory(), torch.empty(nb, dtype=torch.uint8).pin_memory()
        dv_in, dv_out = torch.empty(nb, dtype=torch.uint8, device=dev), torch.empty(nb, dtype=torch.uint8, device=dev)
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        barrier()
        reps = 8
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for it in range(reps + 1):
            if it == 1:
                ev[0].record(s_in); ev[2].record(s_out)
            with torch.cuda.stream(s_in):
                dv_in.copy_(hp_in, non_blocking=True)
            with torch.cuda.stream(s_out):
                hp_out.copy_(dv_out, non_blocking=True)
        ev[1].record(s_in); ev[3].record(s_out)
        torch.cuda.synchronize()
        bw = torch.tensor([reps * nb / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9, reps * nb / (ev[2].elapsed_time(ev[3]) * 1e-3) / 1e9],
                          dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(bw, op=dist.ReduceOp.MIN)
        need = (e2e["h2d_bytes_per_step"] + e2e["d2h_bytes_per_step"]) * (e2e["value"] / world / B) / 1e9 if e2e else None
        host_link = {"h2d_gbs_slowest_rank": float(bw[0].item()), "d2h_gbs_slowest_rank": float(bw[1].item()), "ranks_copying": world,
                     "e2e_traffic_gbs_per_rank": need,
                     "note": "pinned 64 MB copies in both directions at once on every rank; e2e moves h2d + d2h bytes per step per rank"}

    # ---- two camera streams on this GPU (two contexts, two CUDA streams): the head and tail of one stream's batch are filled by
    # the other stream's LK.  An extra figure; `value` above stays the single-stream number.
    multi = None
    if not a.no_e2e and not a.lean:
        ctxs, strs, outs2 = [ctx], [stream], [outs]
        cB = capi.Context(width=w, height=h, max_batch=B, pixel_step=a.pixel_step, min_vector_size=0.2, seed=2, device=local,
                          flow_engine=engine)
        sB = torch.cuda.Stream(device=dev)
        cB.set_stream(sB.cuda_stream)
        bufs = [torch.empty_like(t_) for t_ in (d_next, d_status, d_keep, d_H, d_nv, d_inl, d_mask)]
        oB = capi.MdOutputs(bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), bufs[3].data_ptr(), bufs[4].data_ptr(),
                            bufs[5].data_ptr(), bufs[6].data_ptr(), w, w * h)
        ctxs.append(cB); strs.append(sB); outs2.append(oB)

        def step2(i):
            for c_, o_ in zip(ctxs, outs2):
                c_.raw_process_batch(sets[i % R].data_ptr(), 1, w, frame_bytes, B + 1, False, o_, capi.MD_MEM_DEVICE)

        for i in range(max(a.warmup, 1)):
            step2(i)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in strs]
        for (e0_, _), s_ in zip(ev, strs):
            e0_.record(s_)
        n2 = max(3, a.steps // 2)
        for i in range(n2):
            step2(i)
        for (_, e1_), s_ in zip(ev, strs):
            e1_.record(s_)
        barrier()
        # both streams start together; the job ends when the slower one does
        ms2 = streams.max_over_ranks(max(ev[0][0].elapsed_time(ev[0][1]), ev[0][0].elapsed_time(ev[1][1])))
        multi = {"streams_per_gpu": 2, "value": world * 2 * B * n2 / (ms2 * 1e-3), "unit": UNIT}
        cB.close()

    # ---- per-stream statistics gathered over NCCL (the only collective; off the frame path)
    st = ctx.stats()
    gathered = streams.gather_stats({k: st[k] for k in ("pairs", "mask_pixels", "tracked", "inliers")})

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference's OpenCV chain via cv2, all host threads
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle import cv_chain
        n = 0
        if a.flow_engine == "varflow":
            from oracle import oracle as O

            def cpu_pair(b):
                O.process_pair_varflow(frames_np[b], frames_np[b + 1], pixel_step=a.pixel_step, min_vector_size=0.2, seed=1 + b)
            cores, how = 1, "plain-C oracle: VarFlow::CalcFlow restatement (serial Gauss-Seidel, 1 thread) + chain"
        else:
            def cpu_pair(b):
                cv_chain.process_pair(frames_np[b], frames_np[b + 1], pixel_step=a.pixel_step, seed=1 + b)
            cores = cv_chain.threads()
            how = "cv2 OpenCV routines + oracle egomotion fit" if cv_chain.have_cv2() else "plain-C oracle"
            cpu_pair(0)
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < a.cpu_baseline_seconds and n < 2000:
            cpu_pair(n % B)
            n += 1
        dt = time.perf_counter() - t0
        cpu = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d pairs of the same workload in %.1f s (%s)" % (n, dt, how)}
        if a.flow_engine != "varflow" and cv_chain.have_cv2() and cores > 1 and not a.lean:
            # the same chain on ONE host thread beside the all-threads figure (short sample)
            import cv2
            cv2.setNumThreads(1)
            n1, t1 = 0, time.perf_counter()
            while time.perf_counter() - t1 < min(4.0, a.cpu_baseline_seconds) and n1 < 200:
                cpu_pair(n1 % B)
                n1 += 1
            cpu["value_1_thread"] = n1 / (time.perf_counter() - t1)
            cv2.setNumThreads(cores)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/f32/f64 (integer samples, f32 sums, f64 homography)", "data": "synthetic",
            "config": config_dict(a, world),
            "mpx_per_s": value * N / 1e6,
            "roofline": roofline, "lk_work": lk_taps, "stages": stages, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "stream_stats": gathered, "two_streams_per_gpu": multi, "host_pinning": pinned, "host_link": host_link,
            "e2e_packed_mask": e2e_packed, "e2e_one_context": e2e_one, "graph_replays": st["graph_replays"],
            "ms_per_step_by_rank": [m / a.steps for m in ms_ranks], "priming_steps": priming,
        }
        assert line["config"]["grid_points"] == P
    ctx.close()
    if rank == 0:
        if world == 1 and not a.no_secondary and not vf and (w, h, a.pixel_step) == (1920, 1080, 10):
            line["secondary"] = run_secondary(a)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
