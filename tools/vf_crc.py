#!/usr/bin/env python
"""CRC32 of VarFlow's U, V for fixed inputs: a bit-level regression check when the Gauss-Seidel kernel is restructured."""
import os
import sys
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motion_detection_b200 import capi, synth

for (w, h, lit) in ((333, 211, 1), (640, 480, 1), (640, 480, 0), (1920, 1080, 1)):
    fr, _ = synth.sequence(w, h, 2, seed=1234, blobs=2)
    ctx = capi.Context(width=w, height=h, vf_literal=lit)
    U, V = ctx.varflow(fr[0], fr[1])
    print(w, h, lit, "%08x %08x" % (zlib.crc32(U.tobytes()), zlib.crc32(V.tobytes())))
    ctx.close()
