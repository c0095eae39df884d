import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
from srba_stereo_slam_b200 import frontend, synth
B=int(os.environ.get("MB_PAIRS","128")); W,H=1241,376
frames=synth.stereo_batch(B,W,H,seed=0,unique=8)
fe=frontend.StereoFrontEnd(width=W,height=H,nlevels=8,nfeatures=1000,max_pairs=B,use_graph=0)
fe.upload(frames)
for _ in range(3):
    fe.process_device(B); fe.synchronize()
fe.close()
