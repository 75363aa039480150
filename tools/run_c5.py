#!/usr/bin/env python
"""BASELINE configs[4] / SURVEY 8d "C5": a long 1080p sequence through the node's frame loop, end to end, without ROS.

10 000 frames (default) of the C2 scene (global affine camera motion + moving discs) are generated on the fly from the seed
by producer threads (cv2.warpAffine of the canvas -- input synthesis, not the product), handed over as rgb8 host images like
image_transport does (node.cpp:271), and pushed through motion_detection_b200.node.LiveNode (= imageCallback: skip_frames,
window of F frames, trajectories -> fitSubspace -> clusterEuclidean -> boxes -> CSV rows).  Prints one JSON line: frames/s
end to end, per-1000-frame throughput (drift), device memory before / after (leaks), callbacks, boxes, log rows.

  python tools/run_c5.py --frames 10000 --skip-frames 1 --num-motions 2
"""
import argparse
import json
import os
import queue
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--pixel-step", type=int, default=10)
    ap.add_argument("--skip-frames", type=int, default=1)
    ap.add_argument("--num-motions", type=int, default=2)
    ap.add_argument("--producers", type=int, default=12)
    ap.add_argument("--seed", type=int, default=1234)
    a = ap.parse_args()

    import cv2
    import torch
    from motion_detection_b200 import synth
    from motion_detection_b200.node import LiveNode

    w, h = a.width, a.height
    margin = 384
    canvas = synth.texture(w, h, a.seed, margin=margin).astype(np.float32)
    rng = np.random.default_rng(a.seed + 7919)
    discs = [(rng.uniform(0.2 * w, 0.8 * w), rng.uniform(0.2 * h, 0.8 * h), r, *(rng.uniform(2, 4) * np.array([np.cos(t), np.sin(t)])))
             for r, t in zip((20, 30, 40), rng.uniform(0, 2 * np.pi, 3))]
    period = 100          # the camera swings back and forth so that the view stays on the canvas

    def make(k):
        kk = k % (2 * period)
        kk = kk if kk < period else 2 * period - kk
        M = synth.camera_matrix(w, h, kk)
        Mi = np.linalg.inv(M)
        Mi[:2, 2] += margin
        img = cv2.warpAffine(canvas, Mi[:2].astype(np.float64), (w, h), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP)
        g = np.clip(np.rint(img), 0, 255).astype(np.uint8)
        for (cx, cy, r, vx, vy) in discs:
            x = (cx + vx * k) % w
            y = (cy + vy * k) % h
            cv2.circle(g, (int(round(x)), int(round(y))), int(r), 255, -1)
        return cv2.cvtColor(g, cv2.COLOR_GRAY2RGB)

    q = {}
    cond = threading.Condition()
    window = 64

    def producer(pid):
        for k in range(pid, a.frames, a.producers):
            with cond:
                while k - state["next"] >= window:
                    cond.wait()
            f = make(k)
            with cond:
                q[k] = f
                cond.notify_all()

    state = {"next": 0}
    threads = [threading.Thread(target=producer, args=(i,), daemon=True) for i in range(a.producers)]
    for t in threads:
        t.start()

    node = LiveNode(w, h, pixel_step=a.pixel_step, num_motions=a.num_motions, skip_frames=a.skip_frames, sigma=0.5,
                    distance_threshold=50.0, seed=1)
    free0, _ = torch.cuda.mem_get_info(0)
    marks, boxes, trajs, t_wait = [], 0, 0, 0.0
    t0 = time.perf_counter()
    tm = t0
    for k in range(a.frames):
        tw = time.perf_counter()
        with cond:
            while k not in q:
                cond.wait()
            f = q.pop(k)
            state["next"] = k + 1
            cond.notify_all()
        t_wait += time.perf_counter() - tw
        res = node.on_image(f)
        if res is not None:
            boxes += len(res["boxes"])
            trajs += res["num_trajectories"]
        if (k + 1) % 1000 == 0:
            now = time.perf_counter()
            marks.append(round(1000 / (now - tm), 1))
            tm = now
            if k + 1 == 1000:
                free1, _ = torch.cuda.mem_get_info(0)
    dt = time.perf_counter() - t0
    free2, _ = torch.cuda.mem_get_info(0)
    line = {"workload": "C5: %d frames %dx%d rgb8 through LiveNode (imageCallback without ROS), pixel_step %d, skip_frames %d, F = %d" % (
                a.frames, w, h, a.pixel_step, a.skip_frames, 2 * a.num_motions + 1),
            "frames_per_s_end_to_end": a.frames / dt, "seconds": dt, "seconds_waiting_for_input": t_wait,
            "frames_per_s_excluding_input_wait": a.frames / max(dt - t_wait, 1e-9),
            "frames_per_s_per_1000": marks, "callbacks": node.callbacks, "mean_trajectories": trajs / max(node.callbacks, 1),
            "boxes": boxes, "log_rows": len(node.log_rows), "global_frame_count": node.global_frame_count,
            "device_memory_drift_bytes_after_first_1000": int(free1 - free2) if a.frames >= 1000 else None,
            "device_memory_used_bytes": int(free0 - free2), "producers": a.producers}
    print(json.dumps(line), flush=True)
    node.close()


if __name__ == "__main__":
    main()
