#!/usr/bin/env python
"""Summarise an MD_VF_TRACE dump (k_vf_gs per-warp phase clocks of the finest-level n1+n2 sweep call).
usage: MD_VF_TRACE=/tmp/t.bin python tools/vf_one.py; python tools/vf_trace_summary.py /tmp/t.bin"""
import sys
import numpy as np

a = np.fromfile(sys.argv[1], dtype=np.int64)
steps, warps = int(a[0]), int(a[1])
r = a[2:2 + steps * warps * 8].reshape(steps, warps, 8)
tasks, stage, inner, wb, fence, bar, total = [r[:, :, i] for i in range(7)]
busy = tasks > 0
print("steps %d warps %d; tasks per step: mean %.1f max %d" % (steps, warps, busy.sum(1).mean(), busy.sum(1).max()))
print("per time step (cycles): total mean %.0f  (max over warps, mean over steps %.0f)" % (total.mean(), total.max(1).mean()))
for name, v in (("stage", stage), ("inner", inner), ("write-back", wb)):
    print("  %-10s busy warps: mean %.0f  max-per-step mean %.0f" % (name, v[busy].mean(), v.max(1).mean()))
print("  fence mean %.0f   barrier wait: busy warps mean %.0f, idle warps mean %.0f" % (fence.mean(), bar[busy].mean(), bar[~busy].mean()))
print("  inner per pixel step (63 steps): %.0f cycles" % (inner[busy].mean() / 63))
