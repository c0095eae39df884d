import sys, time, json
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from srba_stereo_slam_b200 import frontend, synth
W,H=1241,376
res={}
for nctx,B in ((1,128),(2,64),(2,128),(3,128),(4,64)):
    frames = synth.stereo_batch(B, W, H, seed=0, unique=8)
    fes=[frontend.StereoFrontEnd(width=W,height=H,nlevels=8,nfeatures=1000,max_pairs=B) for _ in range(nctx)]
    for f in fes: f.upload(frames)
    for _ in range(3):
        for f in fes: f.process_device(B)
    for f in fes: f.synchronize()
    steps=100
    t0=time.perf_counter()
    for _ in range(steps):
        for f in fes: f.process_device(B)
    for f in fes: f.synchronize()
    dt=time.perf_counter()-t0
    res[f"{nctx}x{B}"]=nctx*B*steps/dt
    for f in fes: f.close()
print(json.dumps(res))
