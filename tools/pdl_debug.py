"""PDL A/B: results of a 2-pair batch (graph and direct launches) and the un-batched latency, with the library's SFE_NO_PDL / SFE_PDL_MASK knobs."""
import os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
import numpy as np
from srba_stereo_slam_b200 import frontend, synth
if os.environ.get('LIB'): frontend.LIB_PATH = os.path.join(os.environ.get('GRAFT_REPO_ROOT', '.'), os.environ['LIB'])
frames = synth.stereo_batch(2, 1241, 376, seed=0, unique=2)
fe = frontend.StereoFrontEnd(width=1241, height=376, nlevels=8, nfeatures=1000, max_pairs=2, use_graph=int(os.environ.get("G", "1")))
for it in range(2):
    fe.process_batch_host(frames)
    out = []
    for p in range(2):
        kl, kr, dl, dr, m = fe.get_values(p); out.append((len(kl), len(kr), len(m)))
print("no_pdl", os.environ.get("SFE_NO_PDL"), "mask", os.environ.get("SFE_PDL_MASK"), "graph", os.environ.get("G", "1"), out)
fe.close()
fe = frontend.StereoFrontEnd(width=1241, height=376, nlevels=8, nfeatures=1000, max_pairs=1)
ts = []
for i in range(300):
    t0 = time.perf_counter(); fe.process_stereo(frames[i & 1, 0], frames[i & 1, 1]); fe.get_values(0); ts.append(time.perf_counter() - t0)
print("single pair ms median", float(np.median(ts[50:])) * 1e3)
fe.close()
