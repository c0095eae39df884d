#!/usr/bin/env python
"""Generate tests/golden/*.npz with cv2 (4.13 in the build container).  TEST INFRASTRUCTURE.

The reference's arithmetic lives in OpenCV 2.4 / DBoW2 / stereo-vo, none of which is present here
(SURVEY.md 8(c)).  cv2 4.13's building blocks that are bit-identical to 2.4 by construction are used to
pin the oracle: resize(INTER_LINEAR) with IPP off, FastFeatureDetector, fastAtan2, BFMatcher(NORM_HAMMING),
and -- for nlevels=1, where no pyramid resize is involved -- cv2.ORB end to end (keypoints, Harris
responses, angles and descriptors; its blur is the float sepFilter2D path = our blur_mode 1).
Run from the repo root:  python oracle/tools/make_golden.py
"""
import os, sys
import numpy as np
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from srba_stereo_slam_b200 import synth, voc_loader   # scene renderer / loader are product-side helpers
OUT = os.path.join(ROOT, "tests", "golden")

cv2.ipp.setUseIPP(False)
cv2.setNumThreads(1)


def kp_array(kps):
    return np.array([(k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id) for k in kps], np.float64).reshape(-1, 7)


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261017)
    sc = synth.CorridorScene(seed=3, width=640, height=192)
    l0, r0 = sc.frame(0)
    l5, _ = sc.frame(5)
    noise = cv2.GaussianBlur((rng.random((150, 200)) * 255).astype(np.uint8), (0, 0), 1.2)

    # 1. resize: pyramid-style chains and odd sizes
    cases = []
    for name, img, (dw, dh) in (("scene_533x160", l0, (533, 160)), ("scene_444x133", cv2.resize(l0, (533, 160), interpolation=cv2.INTER_LINEAR), (444, 133)),
                                 ("noise_167x125", noise, (167, 125)), ("noise_199x149", noise, (199, 149)), ("noise_100x75", noise, (100, 75))):
        cases.append((name, img, cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)))
    np.savez_compressed(os.path.join(OUT, "resize.npz"), names=np.array([c[0] for c in cases]),
                        **{"src_%d" % i: c[1] for i, c in enumerate(cases)}, **{"dst_%d" % i: c[2] for i, c in enumerate(cases)})

    # 2. FAST-9/16 + NMS
    d = {}
    for i, (img, th) in enumerate(((l0, 20), (l0, 7), (r0, 30), (noise, 12))):
        kp = cv2.FastFeatureDetector_create(th, True).detect(img)
        d["img_%d" % i] = img; d["th_%d" % i] = th
        d["xys_%d" % i] = np.array([(k.pt[0], k.pt[1], k.response) for k in kp], np.int32).reshape(-1, 3)
    np.savez_compressed(os.path.join(OUT, "fast.npz"), n=4, **d)

    # 3. fastAtan2
    yx = np.concatenate([rng.integers(-200000, 200000, (4000, 2)), [[0, 0], [0, 5], [5, 0], [-5, 0], [0, -5], [7, 7], [-7, 7]]]).astype(np.float32)
    ang = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
    np.savez_compressed(os.path.join(OUT, "atan2.npz"), yx=yx, angle=ang)

    # 4. cv2.ORB end to end at nlevels=1 (the reference's shipped n_levels, UTL.h:288)
    d = {}
    for i, (img, nf, th) in enumerate(((l0, 300, 20), (r0, 500, 20), (l5, 200, 9), (noise, 60, 25))):
        orb = cv2.ORB_create(nfeatures=nf, nlevels=1, fastThreshold=th)
        kp, desc = orb.detectAndCompute(img, None)
        d["img_%d" % i] = img; d["nf_%d" % i] = nf; d["th_%d" % i] = th
        d["kp_%d" % i] = kp_array(kp); d["desc_%d" % i] = desc if desc is not None else np.zeros((0, 32), np.uint8)
    np.savez_compressed(os.path.join(OUT, "orb_nl1.npz"), n=4, **d)

    # 4b. cv2.ORB end to end with the full pyramid (cv4 arithmetic: INTER_LINEAR_EXACT pyramid + float blur):
    #     pins the multi-level logic (level sizes, quotas, per-level selection, coordinate scaling)
    d = {}
    full = synth.CorridorScene(seed=0, width=1241, height=376).frame(1)[0]
    for i, (img, nf, nl, th) in enumerate(((full, 1000, 8, 20), (l0, 300, 4, 20), (r0, 250, 3, 12))):
        orb = cv2.ORB_create(nfeatures=nf, nlevels=nl, fastThreshold=th)
        kp, desc = orb.detectAndCompute(img, None)
        d["img_%d" % i] = img; d["nf_%d" % i] = nf; d["nl_%d" % i] = nl; d["th_%d" % i] = th
        d["kp_%d" % i] = kp_array(kp); d["desc_%d" % i] = desc
    np.savez_compressed(os.path.join(OUT, "orb_multilevel.npz"), n=3, **d)

    # 4c. cv2.ORB end to end off the default parameters: scale factors 1.05 .. 1.25, edge thresholds 19 .. 45, other
    #     FAST thresholds and level counts (cv4 arithmetic) - pins level sizes, quotas and coordinate scaling in general
    d = {}
    mid = synth.CorridorScene(seed=3, width=640, height=400).frame(2)[0]
    d["img"] = mid
    combos = ((400, 6, 1.1, 31, 20), (500, 8, 1.25, 31, 15), (300, 5, 1.2, 19, 20), (300, 4, 1.15, 45, 30), (800, 8, 1.05, 31, 20))
    for i, (nf, nl, sf, et, th) in enumerate(combos):
        orb = cv2.ORB_create(nfeatures=nf, nlevels=nl, scaleFactor=sf, edgeThreshold=et, fastThreshold=th)
        kp, desc = orb.detectAndCompute(mid, None)
        d["prm_%d" % i] = np.array([nf, nl, sf, et, th], np.float64)
        d["kp_%d" % i] = kp_array(kp); d["desc_%d" % i] = desc
    np.savez_compressed(os.path.join(OUT, "orb_params.npz"), n=len(combos), **d)

    # 5. BFMatcher(NORM_HAMMING): random descriptors with planted near-duplicates and exact ties
    Q = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    T = rng.integers(0, 256, (257, 32), dtype=np.uint8)
    T[10] = Q[3]; T[200] = Q[3]                      # exact tie -> lowest train index
    T[50] = Q[7]; T[50, 0] ^= 1; T[51] = Q[7]; T[51, 5] ^= 16   # distance-1 tie
    m = cv2.BFMatcher(cv2.NORM_HAMMING, False).match(Q, T)
    np.savez_compressed(os.path.join(OUT, "bf.npz"), Q=Q, T=T,
                        m=np.array([(x.queryIdx, x.trainIdx, x.imgIdx, x.distance) for x in m], np.float64))

    # 6. the vocabulary, flattened (data derived from the reference's demo/voc.yml.gz, its only golden fixture)
    ref_voc = "/root/reference/demo/voc.yml.gz"
    if os.path.exists(ref_voc):
        v = voc_loader.load_vocabulary(ref_voc)
        np.savez_compressed(os.path.join(OUT, "voc_flat.npz"), **{k: np.asarray(val) for k, val in v.items()})
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  %-16s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    main()
