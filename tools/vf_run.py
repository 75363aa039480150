import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motion_detection_b200 import capi, synth
w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
fr, _ = synth.sequence(w, h, 2, seed=1234, camera=False, blobs=0, patch=True)
ctx = capi.Context(width=w, height=h, vf_literal=0)
ctx.varflow(fr[0], fr[1]); ctx.varflow(fr[0], fr[1])
