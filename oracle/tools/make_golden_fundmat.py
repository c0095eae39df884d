#!/usr/bin/env python
"""Golden vectors for the F-matrix data-association filter: cv2.findFundamentalMat(p1, p2, FM_RANSAC, param1, param2)
inlier masks from THIS container's cv2 (4.13) on seeded two-view scenes with planted outliers.  OpenCV 2.4 (what the
reference links) and 4.13 share the RNG (seed -1), the `rng % count` sampling and the 7-point minimal solver, so the
masks pin the restatement in oracle/geometry_oracle.py and the product's host/fundamental_ransac.cpp.
Writes tests/golden/fundmat.npz.  Run: python oracle/tools/make_golden_fundmat.py"""
import os
import numpy as np
import cv2

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden")


def rodrigues(w):
    return cv2.Rodrigues(np.asarray(w, np.float64))[0]


def main():
    rng = np.random.default_rng(20261017)
    K = np.array([[707.0912, 0, 601.8873], [0, 707.0912, 183.1104], [0, 0, 1]])
    d = {}; n_cases = 40
    for c in range(n_cases):
        n = int(rng.integers(15, 320))
        X = np.c_[rng.uniform(-12, 12, n), rng.uniform(-3, 3, n), rng.uniform(4, 45, n)]
        R = rodrigues([rng.normal(0, 0.01), rng.uniform(-0.15, 0.15), rng.normal(0, 0.01)])
        t = np.array([rng.uniform(-0.4, 0.4), rng.normal(0, 0.02), rng.uniform(0.3, 2.5)])
        p1 = (K @ X.T).T; p1 = p1[:, :2] / p1[:, 2:]
        X2 = (R @ X.T).T + t; p2 = (K @ X2.T).T; p2 = p2[:, :2] / p2[:, 2:]
        noise = rng.choice([0.0, 0.3, 0.7, 1.2])
        p1 = p1 + rng.normal(0, noise, p1.shape) if noise else p1
        p2 = p2 + rng.normal(0, noise, p2.shape) if noise else p2
        no = int(n * rng.uniform(0.0, 0.45)); idx = rng.permutation(n)[:no]
        p2[idx] += rng.uniform(-90, 90, (no, 2))
        if c % 7 == 3:                                    # keypoint-like: quantised to the level-0 pixel grid
            p1 = np.rint(p1); p2 = np.rint(p2)
        p1 = np.ascontiguousarray(p1, np.float32); p2 = np.ascontiguousarray(p2, np.float32)
        th, pr = float(rng.choice([1.0, 1.5, 2.0, 3.0])), float(rng.choice([0.9, 0.95, 0.99]))
        F, m = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, th, pr)
        d["p1_%d" % c] = p1; d["p2_%d" % c] = p2; d["prm_%d" % c] = np.array([th, pr])
        d["mask_%d" % c] = (m.ravel().astype(np.uint8) if m is not None else np.zeros(n, np.uint8))
        d["F_%d" % c] = F if F is not None and F.shape == (3, 3) else np.zeros((3, 3))
    np.savez_compressed(os.path.join(OUT, "fundmat.npz"), n=n_cases, cv2_version=np.array(cv2.__version__), **d)
    print("wrote fundmat.npz:", n_cases, "cases")


if __name__ == "__main__":
    main()
