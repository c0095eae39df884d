"""Small driver for ncu captures: a few direct-launch (no CUDA graph) batches of the headline workload on one context.
MB_PAIRS (default 128) pairs per batch; MB_FULL=1 also runs the keyframe-database batch (BoW + association) with a pre-filled store."""
import os
import sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from srba_stereo_slam_b200 import frontend, synth
B = int(os.environ.get("MB_PAIRS", "128")); W, H = 1241, 376
frames = synth.stereo_batch(B, W, H, seed=0, unique=8)
fe = frontend.StereoFrontEnd(width=W, height=H, nlevels=8, nfeatures=1000, max_pairs=B, use_graph=0)
fe.upload(frames)
full = os.environ.get("MB_FULL") == "1"
if full:
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    vz = np.load(os.path.join(root, "tests", "golden", "voc_flat.npz"))
    voc = {k: (vz[k] if vz[k].ndim else vz[k].item()) for k in vz.files}
    bow = frontend.BowDatabase(fe, voc); db = frontend.KeyframeDatabase(fe, bow, 16 * B)
for i in range(int(os.environ.get("MB_BATCHES", "3"))):
    fe.process_device(B)
    if full:
        db.process_batch(B)
    fe.synchronize()
if full:
    db.close(); bow.close()
fe.close()
