import sys, numpy as np
sys.path.insert(0, '.')
from motion_detection_b200 import capi, synth
w, h = 1920, 1080
frames, _ = synth.sequence(w, h, 2, seed=3)
ctx = capi.Context(width=w, height=h, max_batch=1)
for _ in range(2):
    U, V = ctx.varflow(frames[0], frames[1])
print(float(np.abs(U).mean()))
