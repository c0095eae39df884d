"""Full front-end step (stereo extraction + keyframe database batch) with the database growing 384 -> 4480 keyframes, one context,
host wall clock between synchronisations; LIB=<path> selects another build of the library (A/B).  Prints ms per 128-frame batch and a
digest of every batch's query hits / scores / association keys."""
import hashlib, os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from srba_stereo_slam_b200 import frontend, synth
ROOT = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("LIB"):
    frontend.LIB_PATH = os.path.join(ROOT, os.environ["LIB"])
B, W, H = 128, 1241, 376
frames = synth.stereo_batch(B, W, H, seed=0, unique=64)
if os.environ.get("KFB_NOISE"):
    # independent random textures (blurred noise, right = left shifted by a disparity): frames share few vocabulary words
    rng = np.random.default_rng(7)
    for i in range(B):
        t = rng.integers(0, 256, (H // 4 + 2, W // 4 + 8), dtype=np.uint8).repeat(4, 0).repeat(4, 1).astype(np.float32)
        t = (t + np.roll(t, 1, 0) + np.roll(t, 1, 1) + np.roll(t, 2, 0)) / 4
        frames[i, 0] = t[:H, 16:16 + W].astype(np.uint8); frames[i, 1] = t[:H, 0:W].astype(np.uint8)
fe = frontend.StereoFrontEnd(width=W, height=H, nlevels=8, nfeatures=1000, max_pairs=B)
vz = np.load(os.path.join(ROOT, "tests", "golden", "voc_flat.npz"))
voc = {k: (vz[k] if vz[k].ndim else vz[k].item()) for k in vz.files}
bow = frontend.BowDatabase(fe, voc); db = frontend.KeyframeDatabase(fe, bow, 36 * B)
fe.upload(frames)
h = hashlib.sha256()
for i in range(3):
    fe.process_device(B); db.process_batch(B)
fe.synchronize()
t0 = time.perf_counter()
for i in range(32):
    fe.process_device(B); db.process_batch(B)
fe.synchronize()
ms = (time.perf_counter() - t0) * 1e3 / 32
r = db.results(B)
for k in sorted(r):
    h.update(np.ascontiguousarray(r[k]).tobytes())
print("KF_BENCH lib %s  ms/batch %.4f  frames/s %.0f  db %d  digest %s" % (os.environ.get("LIB", "default"), ms, B / ms * 1e3, db.size(), h.hexdigest()[:16]))
db.close(); bow.close(); fe.close()
