#!/usr/bin/env python
"""Timing of the engines next to the main chain (one JSON line each): VarFlow dense flow, fitSubspace RANSAC,
trajectory tracking.  GPU numbers through the C ABI with host buffers (synchronous calls, wall clock, best of N after
a warm-up); CPU numbers from the oracle (plain C, the serial algorithms the reference runs on one thread).
usage: python tools/bench_engines.py [--cpu]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motion_detection_b200 import capi, synth  # noqa: E402


def best(fn, n=5):
    fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    cpu = "--cpu" in sys.argv
    if cpu:
        from oracle import oracle as O
    out = []
    for (w, h) in ((640, 480), (1920, 1080), (3840, 2160)):
        fr, _ = synth.sequence(w, h, 2, seed=1234, camera=False, blobs=0, patch=True)
        for literal in (1, 0):
            ctx = capi.Context(width=w, height=h, vf_literal=literal)
            t = best(lambda: ctx.varflow(fr[0], fr[1]), 5)
            rec = {"engine": "VarFlow::CalcFlow", "size": [w, h], "literal_corrections": literal, "gpu_ms": 1e3 * t,
                   "mpx_per_s": w * h / t / 1e6, "algorithmic_bytes": 10 * w * h, "gbs": 10 * w * h / t / 1e9}
            if cpu and literal == 1 and w <= 1920:
                rec["cpu_ms_oracle_1thread"] = 1e3 * best(lambda: O.varflow(fr[0], fr[1]), 2)
            ctx.close()
            out.append(rec)
    for T in (3000, 20736, 200000):
        traj, _ = synth.trajectories(T, 5, num_motions=2, seed=T, noise=0.05)
        ctx = capi.Context(width=64, height=64)
        t = best(lambda: ctx.fit_subspace(traj, num_motions=2, sigma=0.5, seed=1), 5)
        rec = {"engine": "OutlierDetector::fitSubspace", "trajectories": T, "frames": 5, "gpu_ms": 1e3 * t}
        if cpu:
            rec["cpu_ms_oracle_1thread"] = 1e3 * best(lambda: O.fit_subspace(traj, num_motions=2, sigma=0.5, seed=1), 2)
        ctx.close()
        out.append(rec)
    for (w, h) in ((640, 480), (1920, 1080)):
        fr, _ = synth.sequence(w, h, 5, seed=1234)
        ctx = capi.Context(width=w, height=h, max_batch=4, min_vector_size=0.2)
        t = best(lambda: ctx.track_trajectories(fr), 5)
        rec = {"engine": "calculateOpticalFlowTrajectory (F=5)", "size": [w, h], "points": ctx.P, "gpu_ms": 1e3 * t}
        ctx.close()
        out.append(rec)
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
