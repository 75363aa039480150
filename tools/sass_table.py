#!/usr/bin/env python
"""profiles/sass_tma.txt: Blackwell-specific SASS mnemonics per kernel of the built library.
usage: python tools/sass_table.py > profiles/sass_tma.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "motion_detection_b200", "lib", "libmotion_b200.so")
COLS = ["UTMALDG", "SYNCS", "ELECT", "REDUX", "IDP", "VIMNMX3", "VOTE", "DFMA"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = {}
    for m in re.finditer(r"Function : (\S+)", txt):
        names[m.group(1)] = None
    dem = subprocess.run(["cu++filt"] + list(names), capture_output=True, text=True).stdout.splitlines()
    for k, d in zip(list(names), dem):
        m = re.match(r"(.*>)\(", d)
        names[k] = (m.group(1) if m else d.split("(")[0]).replace("(int)", "").replace("(bool)", "")
    print("# cuobjdump -sass motion_detection_b200/lib/libmotion_b200.so : Blackwell-specific mnemonics per kernel (tools/sass_table.py)")
    print("# UTMALDG = cp.async.bulk.tensor (TMA tile load), SYNCS = mbarrier, ELECT = elect.sync, REDUX = warp reduce, IDP = dp2a / dp4a,")
    print("# VIMNMX3 = three-input min/max, VOTE = ballot, DFMA = f64 fused multiply-add; total = instructions in the kernel")
    print("%-60s" % "kernel" + "".join("%9s" % c for c in COLS) + "    total")
    for blk in txt.split("Function : ")[1:]:
        fn = blk.split()[0]
        ins = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", blk, flags=re.M)
        row = [sum(1 for i in ins if i.split(".")[0] == c) for c in COLS]
        print("%-60s" % names.get(fn, fn)[:60] + "".join("%9d" % v for v in row) + "%9d" % len(ins))


if __name__ == "__main__":
    main()
