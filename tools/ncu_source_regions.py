#!/usr/bin/env python
"""Per-region executed instructions and stall samples of one kernel from `ncu --page source --csv`.
usage: python tools/ncu_source_regions.py source.csv npoints name:start:end ...   (offsets in hex, relative to the kernel start)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
npts = float(sys.argv[2])
h = rows[1]; data = rows[2:]
ia, isamp, iex = h.index('Address'), h.index('# Samples'), h.index('Instructions Executed')
base = int(data[0][ia], 16)
regions = [(a.split(':')[0], int(a.split(':')[1], 16), int(a.split(':')[2], 16)) for a in sys.argv[3:]]
tot_s = sum(int(r[isamp]) for r in data); tot_e = sum(int(r[iex]) for r in data)
print(f'total: {tot_e / npts:.1f} warp-instr / point, {tot_s} samples')
stall = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
for name, a, b in regions:
    sel = [r for r in data if a <= int(r[ia], 16) - base < b]
    s = sum(int(r[isamp]) for r in sel); e = sum(int(r[iex]) for r in sel)
    d = {c: sum(int(r[h.index(c)]) for r in sel) for c in stall}
    t = sum(d.values()) or 1
    top = ' '.join(f'{k[6:]}={100 * v / t:.0f}' for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:5])
    print(f'{name:12s} static {len(sel):4d}  exec {e / npts:8.1f}/pt ({100 * e / tot_e:5.1f}%)  samples {100 * s / tot_s:5.1f}%   {top}')
