"""GPU (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): bit-exact pyramid pixels, FAST keypoint sets, match indices and BoW word
ids; descriptors <= 0.1 % bit disagreement (we assert 0), orientation within 1e-3 rad (we assert bit-equal
degrees).  Field lists follow the reference's equality helpers (srba-stereo-slam_utils.cpp:49-59,75).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    from srba_stereo_slam_b200 import frontend
    from oracle import oracle as orc
    return frontend, orc


def _roi_filter(xs, ys, sc, w, h, e=31):
    m = (xs >= e) & (xs < w - e) & (ys >= e) & (ys < h - e)
    return set(zip(xs[m].tolist(), ys[m].tolist(), sc[m].tolist()))


def _assert_keypoints_equal(a, b):
    assert len(a) == len(b)
    for f in ("x", "y", "size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(a[f], b[f]), f


@pytest.mark.parametrize("blur_mode", [0, 1])
def test_config_b_full_pipeline(mods, scene_frames, blur_mode):
    """1241x376, 1000 ORB, 8 levels, L/R row-band match (BASELINE.json configs[1]) on 2 stereo pairs."""
    frontend, orc = mods
    frames = scene_frames(2)
    fe = frontend.StereoFrontEnd(max_pairs=2, blur_mode=blur_mode)
    try:
        assert fe.process_batch_host(frames) == 2
        for p in range(2):
            kl, kr, dl, dr, m = fe.get_values(p)
            sides = []
            for s, (gk, gd) in enumerate(((kl, dl), (kr, dr))):
                img = frames[p, s]
                ok, od, ex = orc.orb_extract(img, nfeatures=1000, nlevels=8, fast_threshold=20, want_levels=True, blur_mode=blur_mode)
                for l in range(8):
                    assert np.array_equal(fe.get_level(2 * p + s, l), ex["pyramid"][l]), ("pyramid", p, s, l)
                    assert np.array_equal(fe.get_level(2 * p + s, l, blurred=True), ex["blurred"][l]), ("blur", p, s, l)
                    xs, ys, sc = orc.fast9_nms(ex["pyramid"][l], 20)
                    gx, gy, gs = fe.get_raw_corners(2 * p + s, l)
                    h, w = ex["pyramid"][l].shape
                    assert set(zip(gx.tolist(), gy.tolist(), gs.tolist())) == _roi_filter(xs, ys, sc, w, h), ("fast", p, s, l)
                _assert_keypoints_equal(gk, ok)
                assert np.array_equal(gd, od)
                assert len(gk) >= 900
                sides.append((ok, od))
            om = orc.match_rowband(sides[0][0], sides[0][1], sides[1][0], sides[1][1], max_y_diff=2.0, max_dist=60)
            assert len(om) > 100
            for f in ("queryIdx", "trainIdx", "imgIdx", "distance"):
                assert np.array_equal(m[f], om[f]), f
    finally:
        fe.close()


def test_config_a_shipped_ini(mods, scene_frames):
    """kitti.ini as shipped: 1226x370 camera, 500 features, n_levels = 1, FAST 20, ORB 60."""
    frontend, orc = mods
    frames = scene_frames(1, width=1226, height=370, seed=1)
    fe = frontend.StereoFrontEnd(width=1226, height=370, nlevels=1, nfeatures=500, max_pairs=1)
    try:
        fe.process_stereo(frames[0, 0], frames[0, 1])
        kl, kr, dl, dr, m = fe.get_values(0)
        ol, odl, _ = orc.orb_extract(frames[0, 0], nfeatures=500, nlevels=1)
        orr, odr, _ = orc.orb_extract(frames[0, 1], nfeatures=500, nlevels=1)
        _assert_keypoints_equal(kl, ol); _assert_keypoints_equal(kr, orr)
        assert np.array_equal(dl, odl) and np.array_equal(dr, odr)
        om = orc.match_rowband(ol, odl, orr, odr)
        assert np.array_equal(m["trainIdx"], om["trainIdx"]) and np.array_equal(m["queryIdx"], om["queryIdx"])
    finally:
        fe.close()


def test_golden_cv2_orb_single_level(mods, golden):
    """The CUDA path against the committed cv2.ORB(nlevels=1) outputs directly (blur_mode 1)."""
    frontend, _ = mods
    g = golden("orb_nl1.npz")
    for i in range(int(g["n"])):
        img = g["img_%d" % i]
        h, w = img.shape
        fe = frontend.StereoFrontEnd(width=w, height=h, nlevels=1, nfeatures=int(g["nf_%d" % i]),
                                     fast_threshold=int(g["th_%d" % i]), blur_mode=1, max_pairs=1)
        try:
            fe.process_stereo(img, img)
            kl, kr, dl, dr, m = fe.get_values(0)
        finally:
            fe.close()
        ref, rdesc = g["kp_%d" % i], g["desc_%d" % i]
        o = np.lexsort((ref[:, 0], ref[:, 1]))
        ref, rdesc = ref[o], rdesc[o]
        assert len(kl) == len(ref)
        for j, f in enumerate(("x", "y", "size", "angle", "response")):
            assert np.array_equal(kl[f], ref[:, j].astype(np.float32)), f
        assert np.array_equal(dl, rdesc) and np.array_equal(dr, rdesc)
        # identical images: every left keypoint matches itself at distance 0 (disparity 0 is allowed)
        assert np.array_equal(m["queryIdx"], m["trainIdx"]) and np.all(m["distance"] == 0) and len(m) == len(kl)


def test_golden_cv2_orb_multilevel(mods, golden):
    """The CUDA path in cv4 arithmetic (pyramid_mode 1 + blur_mode 1) against the committed cv2.ORB outputs at
    8 / 4 / 3 levels: every keypoint field and descriptor bit identical, end to end."""
    frontend, orc = mods
    g = golden("orb_multilevel.npz")
    for i in range(int(g["n"])):
        img = g["img_%d" % i]
        h, w = img.shape
        fe = frontend.StereoFrontEnd(width=w, height=h, nlevels=int(g["nl_%d" % i]), nfeatures=int(g["nf_%d" % i]),
                                     fast_threshold=int(g["th_%d" % i]), blur_mode=1, pyramid_mode=1, max_pairs=1)
        try:
            fe.process_stereo(img, img)
            kl, kr, dl, dr, m = fe.get_values(0)
        finally:
            fe.close()
        ref, rdesc = g["kp_%d" % i], g["desc_%d" % i]
        sc = np.array([orc.level_scale(int(o)) for o in ref[:, 5]])
        o = np.lexsort((np.rint(ref[:, 0] / sc), np.rint(ref[:, 1] / sc), ref[:, 5]))
        ref, rdesc = ref[o], rdesc[o]
        assert len(kl) == len(ref) > 100
        for j, f in enumerate(("x", "y", "size", "angle", "response", "octave", "class_id")):
            assert np.array_equal(kl[f], ref[:, j].astype(kl[f].dtype)), (i, f)
        assert np.array_equal(dl, rdesc) and np.array_equal(dr, rdesc)


def test_golden_cv2_orb_parameter_variants(mods, golden):
    """The CUDA path against cv2.ORB outputs with scale factors 1.05 .. 1.25, edge thresholds 19 .. 45, 4 .. 8 levels
    (cv4 arithmetic): every keypoint field and descriptor bit, straight against the committed cv2 vectors."""
    frontend, orc = mods
    g = golden("orb_params.npz")
    img = g["img"]
    h, w = img.shape
    for i in range(int(g["n"])):
        nf, nl, sf, et, th = g["prm_%d" % i]
        if not 1.05 <= sf <= 1.25:
            continue
        fe = frontend.StereoFrontEnd(width=w, height=h, nlevels=int(nl), nfeatures=int(nf), scale_factor=float(sf), edge_threshold=int(et),
                                     fast_threshold=int(th), blur_mode=1, pyramid_mode=1, max_pairs=1)
        try:
            fe.process_stereo(img, img)
            kl, kr, dl, dr, m = fe.get_values(0)
        finally:
            fe.close()
        ref, rdesc = g["kp_%d" % i], g["desc_%d" % i]
        sc = np.array([orc.level_scale(int(o), float(sf)) for o in ref[:, 5]])
        o = np.lexsort((np.rint(ref[:, 0] / sc), np.rint(ref[:, 1] / sc), ref[:, 5]))
        ref, rdesc = ref[o], rdesc[o]
        assert len(kl) == len(ref) == int(nf)
        for j, f in enumerate(("x", "y", "size", "angle", "response", "octave", "class_id")):
            assert np.array_equal(kl[f], ref[:, j].astype(kl[f].dtype)), (i, f)
        assert np.array_equal(dl, rdesc) and np.array_equal(dr, rdesc)


def test_runtime_thresholds_and_repeat(mods, scene_frames):
    """setFASTThreshold / setORBThreshold + request.repeat (EST.cpp:263-315)."""
    frontend, orc = mods
    frames = scene_frames(1)
    fe = frontend.StereoFrontEnd(max_pairs=1, nfeatures=600)
    try:
        fe.process_stereo(frames[0, 0], frames[0, 1])
        k20 = fe.get_values(0)[0]
        fe.set_fast_threshold(45); fe.set_orb_threshold(35)
        assert fe.get_fast_threshold() == 45 and fe.get_orb_threshold() == 35
        fe.process_stereo(None, None, repeat=True)
        kl, kr, dl, dr, m = fe.get_values(0)
        ol, odl, _ = orc.orb_extract(frames[0, 0], nfeatures=600, fast_threshold=45)
        orr, odr, _ = orc.orb_extract(frames[0, 1], nfeatures=600, fast_threshold=45)
        _assert_keypoints_equal(kl, ol); _assert_keypoints_equal(kr, orr)
        om = orc.match_rowband(ol, odl, orr, odr, max_dist=35)
        assert np.array_equal(m["trainIdx"], om["trainIdx"]) and np.array_equal(m["distance"], om["distance"])
        assert np.all(m["distance"] <= 35)
        assert len(kl) < len(k20) == 600      # a FAST threshold of 45 starves the quota on this scene
    finally:
        fe.close()


def test_edge_cases(mods):
    frontend, orc = mods
    fe = frontend.StereoFrontEnd(width=320, height=240, nlevels=4, nfeatures=200, max_pairs=1, robust_1to1=1)
    try:
        flat = np.full((240, 320), 128, np.uint8)
        fe.process_stereo(flat, flat)                                   # no corners at all
        kl, kr, dl, dr, m = fe.get_values(0)
        assert len(kl) == len(kr) == len(m) == 0
        rng = np.random.default_rng(5)
        img = (rng.random((240, 320)) * 255).astype(np.uint8)           # white noise: corners everywhere, many score ties
        fe.process_stereo(img, img[:, ::-1].copy())
        kl, kr, dl, dr, m = fe.get_values(0)
        ol, odl, _ = orc.orb_extract(img, nfeatures=200, nlevels=4)
        orr, odr, _ = orc.orb_extract(img[:, ::-1].copy(), nfeatures=200, nlevels=4)
        _assert_keypoints_equal(kl, ol); _assert_keypoints_equal(kr, orr)
        assert np.array_equal(dl, odl) and np.array_equal(dr, odr)
        om = orc.match_rowband(ol, odl, orr, odr, robust_1to1=True)
        assert np.array_equal(m["queryIdx"], om["queryIdx"]) and np.array_equal(m["trainIdx"], om["trainIdx"])
    finally:
        fe.close()
    with pytest.raises(frontend.SfeError):
        frontend.StereoFrontEnd(width=32, height=32)
    with pytest.raises(frontend.SfeError):
        frontend.StereoFrontEnd(nlevels=9)


@pytest.mark.parametrize("blur_mode", [0, 1])
def test_saturated_and_odd_sized_images(mods, blur_mode):
    """Over-exposed blocks (row sums at the 16-bit limit, column sums above 2^24: the blur's saturation path), hard
    black/white edges on the image border (reflect-101 halo) and an odd size that leaves partial tiles on every level."""
    frontend, orc = mods
    W, H = 413, 287
    rng = np.random.default_rng(11)
    img = np.zeros((H, W), np.uint8)
    img[:, :] = (rng.random((H, W)) * 2).astype(np.uint8) + 254              # 254/255 everywhere
    img[40:120, 60:200] = (rng.random((80, 140)) * 255).astype(np.uint8)     # a textured patch
    img[150:, :30] = 0; img[:25, 300:] = 0; img[260:, 380:] = 255            # blocks touching the borders
    img[200:240, 100:300:2] = 0                                              # 1-pixel stripes
    fe = frontend.StereoFrontEnd(width=W, height=H, nlevels=5, nfeatures=300, max_pairs=1, blur_mode=blur_mode)
    try:
        fe.process_stereo(img, img[::-1].copy())
        for s, im in enumerate((img, img[::-1].copy())):
            ok, od, ex = orc.orb_extract(im, nfeatures=300, nlevels=5, fast_threshold=20, want_levels=True, blur_mode=blur_mode)
            for l in range(5):
                assert np.array_equal(fe.get_level(s, l), ex["pyramid"][l]), ("pyramid", s, l)
                assert np.array_equal(fe.get_level(s, l, blurred=True), ex["blurred"][l]), ("blur", s, l)
            assert ex["blurred"][0].max() == 255
            kl, kr, dl, dr, m = fe.get_values(0)
            gk, gd = (kl, dl) if s == 0 else (kr, dr)
            _assert_keypoints_equal(gk, ok)
            assert np.array_equal(gd, od)
    finally:
        fe.close()


def test_match_bf(mods, golden):
    """cv::BFMatcher(NORM_HAMMING,false).match (EST.cpp:1447-1452): golden vectors + oracle at 1000x1000 and ragged sizes."""
    frontend, orc = mods
    fe = frontend.StereoFrontEnd(width=320, height=240, nlevels=1, nfeatures=100, max_pairs=1)
    try:
        g = golden("bf.npz")
        m = fe.match_bf(g["Q"], g["T"])
        ref = g["m"]
        assert np.array_equal(m["queryIdx"], ref[:, 0]) and np.array_equal(m["trainIdx"], ref[:, 1])
        assert np.array_equal(m["distance"], ref[:, 3].astype(np.float32)) and np.all(m["imgIdx"] == 0)
        rng = np.random.default_rng(11)
        for nq, nt in ((1000, 1000), (1, 1), (37, 4000), (4000, 13), (129, 129), (500, 0), (0, 500)):
            Q = rng.integers(0, 256, (nq, 32), dtype=np.uint8); T = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
            if nt > 8 and nq > 8:
                T[nt // 2] = Q[nq // 3]; T[nt - 1] = Q[nq // 3]; T[0] = T[min(5, nt - 1)]    # planted ties
            m = fe.match_bf(Q, T); om = orc.match_bf(Q, T)
            assert len(m) == len(om)
            for f in ("queryIdx", "trainIdx", "imgIdx", "distance"):
                assert np.array_equal(m[f], om[f]), (nq, nt, f)
    finally:
        fe.close()


def test_bow_transform_vector_query(mods, voc, scene_frames):
    """CBoWManager (CBoWManager.h:44-111): word ids bit-exact, FP64 BoW vectors and scores bit-identical."""
    frontend, orc = mods
    fe = frontend.StereoFrontEnd(max_pairs=1)
    db = frontend.BowDatabase(fe, voc)
    try:
        rng = np.random.default_rng(7)
        frames = scene_frames(2)
        sets = []
        for p in range(2):
            k, d, _ = orc.orb_extract(frames[p, 0]); sets.append(d)
        sets.append(rng.integers(0, 256, (1000, 32), dtype=np.uint8))
        sets.append(rng.integers(0, 256, (3, 32), dtype=np.uint8))
        sets.append(np.repeat(sets[0][:5], 50, axis=0))                # repeated words -> accumulation order
        odb = []
        for i, d in enumerate(sets):
            w, wt = db.transform(d); ow, owt = orc.bow_transform(voc, d)
            assert np.array_equal(w, ow) and np.array_equal(wt, owt), i
            vw, vv = db.vector(d); ovw, ovv = orc.bow_vector(voc, ow, owt)
            assert np.array_equal(vw, ovw) and np.array_equal(vv, ovv), i
            assert db.add(d) == i
            odb.append((ovw, ovv))
        assert db.size() == len(sets)
        queries = sets + [np.concatenate([sets[0][:400], sets[1][:300], sets[2][:100]])]
        for q in queries:
            res = db.query(q, 4); ores = orc.bow_query(voc, odb, q, 4)
            assert [r[0] for r in res] == [r[0] for r in ores]
            assert [r[1] for r in res] == [r[1] for r in ores]
        top = db.query(sets[0], 4)[0]
        assert top[0] == 0 and abs(top[1] - 1.0) < 1e-12          # a keyframe is its own best match
    finally:
        db.close(); fe.close()


def test_capacity_and_larger_config(mods, scene_frames):
    """1920x1080, 4000 ORB, 8 levels (BASELINE.json configs[4]), one pair: keypoint sets + descriptors + matches."""
    frontend, orc = mods
    from srba_stereo_slam_b200 import synth
    frames = synth.CorridorScene(seed=2, width=1920, height=1080).sequence(1)
    fe = frontend.StereoFrontEnd(width=1920, height=1080, nfeatures=4000, max_pairs=1)
    try:
        fe.process_batch_host(frames)
        kl, kr, dl, dr, m = fe.get_values(0)
        ol, odl, _ = orc.orb_extract(frames[0, 0], nfeatures=4000)
        orr, odr, _ = orc.orb_extract(frames[0, 1], nfeatures=4000)
        _assert_keypoints_equal(kl, ol); _assert_keypoints_equal(kr, orr)
        assert np.array_equal(dl, odl) and np.array_equal(dr, odr)
        om = orc.match_rowband(ol, odl, orr, odr)
        assert np.array_equal(m["trainIdx"], om["trainIdx"]) and np.array_equal(m["queryIdx"], om["queryIdx"])
    finally:
        fe.close()


def test_full_batch_properties(mods, scene_frames):
    """BASELINE.json configs[1] at the benchmark's batch size (128 pairs resident): size-independent properties --
    determinism (idempotent re-run), replica consistency (tiled copies of a frame give identical results), canonical
    ordering, match constraints, and the self-match identity of the brute-force matcher."""
    frontend, orc = mods
    B = 128
    base = scene_frames(2)
    frames = np.ascontiguousarray(np.tile(base, (B // 2, 1, 1, 1)))
    fe = frontend.StereoFrontEnd(max_pairs=B)
    try:
        fe.upload(frames)
        fe.process_device(B); fe.synchronize()
        first = [fe.get_values(p) for p in (0, 1, 64, 65, 126, 127)]
        fe.process_device(B); fe.synchronize()                      # idempotence
        again = [fe.get_values(p) for p in (0, 1, 64, 65, 126, 127)]
        for a, b in zip(first, again):
            for x, y in zip(a, b):
                assert np.array_equal(x, y)
        for i in (2, 4):                                            # replicas of pair 0 / pair 1
            for x, y in zip(first[0], first[i]):
                assert np.array_equal(x, y)
            for x, y in zip(first[1], first[i + 1]):
                assert np.array_equal(x, y)
        kl, kr, dl, dr, m = first[0]
        ol, odl, _ = orc.orb_extract(base[0, 0])
        assert np.array_equal(kl["x"], ol["x"]) and np.array_equal(dl, odl)      # and still the oracle's answer
        for k in (kl, kr):                                           # canonical order: octave, then y, then x (level coords)
            s = 1.2 ** k["octave"].astype(np.float64)
            key = k["octave"].astype(np.int64) * 10**8 + np.rint(k["y"] / s).astype(np.int64) * 10**4 + np.rint(k["x"] / s).astype(np.int64)
            assert np.all(np.diff(key) > 0)
        assert np.all(np.diff(m["queryIdx"]) > 0) and np.all(m["distance"] <= 60) and np.all(m["imgIdx"] == 0)
        assert np.all(np.abs(kl["y"][m["queryIdx"]] - kr["y"][m["trainIdx"]]) <= 2.0)
        assert np.all(kl["x"][m["queryIdx"]] - kr["x"][m["trainIdx"]] >= 0.0)
        self_m = fe.match_bf(dl, dl)                                 # identity, unless two descriptors coincide
        assert np.all(self_m["distance"] == 0) and np.all(self_m["trainIdx"] <= self_m["queryIdx"])
    finally:
        fe.close()


def test_abi_error_paths(mods, scene_frames):
    """Integer status codes across the ABI (no exceptions): call-order and argument errors are reported, not UB."""
    frontend, _ = mods
    fe = frontend.StereoFrontEnd(width=320, height=240, nlevels=2, nfeatures=100, max_pairs=2)
    try:
        with pytest.raises(frontend.SfeError) as e:
            fe.get_values(0)                                            # nothing processed yet
        assert e.value.code == frontend.SFE_ERR_STATE
        with pytest.raises(frontend.SfeError) as e:
            fe.process_stereo(None, None, repeat=True)                  # repeat without a previous frame
        assert e.value.code == frontend.SFE_ERR_STATE
        with pytest.raises(frontend.SfeError) as e:
            fe.process_device(3)                                        # more pairs than the context was sized for
        assert e.value.code == frontend.SFE_ERR_INVALID
        with pytest.raises(frontend.SfeError):
            fe.set_fast_threshold(0)
        rng = np.random.default_rng(0)
        img = (rng.random((2, 240, 320)) * 255).astype(np.uint8)
        fe.process_stereo(img[0], img[1])
        with pytest.raises(frontend.SfeError) as e:
            fe.get_values(1)                                            # only pair 0 exists
        assert e.value.code == frontend.SFE_ERR_INVALID
        kl = fe.get_values(0)[0]
        assert len(kl) > 0                                              # the context is still usable after the errors
        assert len(fe.match_bf(np.zeros((0, 32), np.uint8), np.zeros((5, 32), np.uint8))) == 0
    finally:
        fe.close()


def test_strided_and_ragged_batches(mods, scene_frames):
    """Host images with row/image strides (the generic pitched upload path) and a batch smaller than max_pairs give
    the same results as the tightly packed full path."""
    import ctypes as C
    frontend, _ = mods
    frames = scene_frames(2, width=640, height=192, seed=3)
    W, H = 640, 192
    fe = frontend.StereoFrontEnd(width=W, height=H, nlevels=4, nfeatures=300, max_pairs=8)
    try:
        fe.process_batch_host(frames)                                   # packed, 2 of 8 pairs
        ref = [fe.get_values(p) for p in range(2)]
        stride, istride = 704, 704 * (H + 5)                            # padded rows and padded images
        buf = np.full(4 * istride, 255, np.uint8)
        for i in range(4):
            view = buf[i * istride:i * istride + H * stride].reshape(H, stride)
            view[:, :W] = frames.reshape(4, H, W)[i]
        r = fe.lib.sfe_process_batch_host(fe.h, buf.ctypes.data_as(C.c_void_p), 2, stride, istride)
        assert r == frontend.SFE_OK
        for p in range(2):
            for a, b in zip(ref[p], fe.get_values(p)):
                assert np.array_equal(a, b)
        fe.process_stereo(frames[1, 0], frames[1, 1])                   # single-pair entry on the same context
        for a, b in zip(ref[1], fe.get_values(0)):
            assert np.array_equal(a, b)
    finally:
        fe.close()


def test_two_contexts_joined_streams(mods, scene_frames):
    """sfe_stream_wait: two contexts used round-robin, joined on the device, give the results of one context."""
    frontend, _ = mods
    frames = scene_frames(2)
    a = frontend.StereoFrontEnd(max_pairs=2); b = frontend.StereoFrontEnd(max_pairs=2)
    try:
        a.upload(frames); b.upload(frames[::-1].copy())
        a.timer_start()
        for i in range(4):
            (a, b)[i % 2].process_device(2)
        a.stream_wait(b)
        assert a.timer_stop() > 0.0
        b.synchronize()
        ka = a.get_values(0); kb = b.get_values(1)              # pair 0 of a == pair 1 of b (b holds the batch reversed)
        for x, y in zip(ka, kb):
            assert np.array_equal(x, y)
        a.stream_wait(a)                                        # joining a context with itself is a no-op
    finally:
        a.close(); b.close()


def test_dependent_launch_knob_changes_nothing(mods, scene_frames, monkeypatch):
    """The kernels of a batch are launched programmatically dependent on each other (griddepcontrol.wait before the first access
    to a predecessor's output); SFE_NO_PDL=1 falls back to plain stream order.  Graph replay, direct launches and the un-batched
    call must give the same bytes either way, batch after batch."""
    frontend, _ = mods
    frames = scene_frames(3)
    out = {}
    for knob in ("pdl", "plain"):
        if knob == "plain":
            monkeypatch.setenv("SFE_NO_PDL", "1")            # read when the context is created
        for graph in (1, 0):
            fe = frontend.StereoFrontEnd(max_pairs=3, use_graph=graph)
            try:
                res = []
                for it in range(3):
                    fe.process_batch_host(frames if it != 1 else frames[::-1].copy())
                    res.append([fe.get_values(p) for p in range(3)])
                fe.process_stereo(frames[2, 0], frames[2, 1])
                res.append([fe.get_values(0)])
                out[(knob, graph)] = res
            finally:
                fe.close()
    ref = out[("plain", 0)]
    for key, res in out.items():
        for a_batch, b_batch in zip(res, ref):
            for a_pair, b_pair in zip(a_batch, b_batch):
                assert len(a_pair[0]) > 100
                for x, y in zip(a_pair, b_pair):
                    assert np.array_equal(x, y), key


def test_forked_pyramid_graph_equals_direct_launches(mods, scene_frames, monkeypatch):
    """From 64 images on, the batch graph runs the pyramid phase as two image halves on two streams (fork / join events captured
    into the graph).  The graph replay must give, pair for pair, what plain direct launches on one stream give, with and without
    the fork (SFE_NO_PYR_FORK=1), batch after batch."""
    frontend, _ = mods
    B = 40                                                            # 80 images: above the fork threshold, halves of 40
    base = scene_frames(4)
    frames = np.ascontiguousarray(np.concatenate([base] * (B // 4), axis=0))
    frames[1::2] = frames[1::2, :, ::-1].copy()[:, ::-1]              # every other pair swapped L/R: the halves differ
    ref = frontend.StereoFrontEnd(max_pairs=B, use_graph=0)
    try:
        ref.process_batch_host(frames)
        want = [ref.get_values(p) for p in range(B)]
    finally:
        ref.close()
    assert len(want[0][0]) > 100
    for knob in ("fork", "one stream"):
        if knob == "one stream":
            monkeypatch.setenv("SFE_NO_PYR_FORK", "1")
        fe = frontend.StereoFrontEnd(max_pairs=B, use_graph=1)
        try:
            for it in range(3):
                fe.process_batch_host(frames)
                for p in range(B):
                    for x, y in zip(fe.get_values(p), want[p]):
                        assert np.array_equal(x, y), (knob, it, p)
        finally:
            fe.close()


@pytest.mark.parametrize("w,h,nlevels,scale,nfeat,fth,eth,bm,pm", [
    (401, 303, 1, 1.2, 150, 7, 31, 0, 0),         # single level, low FAST threshold (many quick-test survivors)
    (517, 389, 6, 1.1, 700, 35, 31, 0, 0),        # smaller scale step: wider source footprints in the pyramid kernel
    (640, 480, 8, 1.25, 1200, 20, 19, 1, 1),      # largest scale step, smallest edge threshold, cv4 arithmetic
    (333, 777, 5, 1.2, 400, 60, 40, 1, 0),        # portrait image, high threshold, mixed modes
    (1023, 257, 2, 1.15, 3000, 12, 25, 0, 1),     # wide noisy strip: more keypoints per level than the shared-memory sort holds
    (160, 120, 3, 1.2, 50, 127, 31, 0, 0),        # highest legal FAST threshold: (almost) nothing detected
])
def test_parameter_sweep(mods, w, h, nlevels, scale, nfeat, fth, eth, bm, pm):
    """Configurations off the beaten path of the benchmark: every keypoint field and descriptor bit vs the oracle."""
    frontend, orc = mods
    rng = np.random.default_rng(w * 7 + h)
    yy, xx = np.mgrid[0:h, 0:w]
    img = (96 + 60 * np.sin(xx / 9.0) * np.cos(yy / 7.0)).astype(np.float32)
    for _ in range(60):                                                        # rectangles of random grey: real corners
        x0, y0 = int(rng.integers(0, w - 8)), int(rng.integers(0, h - 8))
        ww, hh = int(rng.integers(4, 60)), int(rng.integers(4, 60))
        img[y0:y0 + hh, x0:x0 + ww] = rng.integers(0, 256)
    img = np.clip(img + rng.normal(0, 30 if nfeat >= 2500 else 3, (h, w)), 0, 255).astype(np.uint8)
    right = np.roll(img, -5, axis=1).copy()
    fe = frontend.StereoFrontEnd(width=w, height=h, nlevels=nlevels, nfeatures=nfeat, scale_factor=scale, fast_threshold=fth,
                                 edge_threshold=eth, blur_mode=bm, pyramid_mode=pm, max_pairs=1)
    try:
        fe.process_stereo(img, right)
        kl, kr, dl, dr, m = fe.get_values(0)
        sides = []
        for im, gk, gd in ((img, kl, dl), (right, kr, dr)):
            ok, od, _ = orc.orb_extract(im, nfeatures=nfeat, nlevels=nlevels, fast_threshold=fth, edge_threshold=eth, scale=scale,
                                        blur_mode=bm, pyramid_mode=pm)
            _assert_keypoints_equal(gk, ok)
            assert np.array_equal(gd, od)
            sides.append((ok, od))
        om = orc.match_rowband(sides[0][0], sides[0][1], sides[1][0], sides[1][1], max_y_diff=2.0, max_dist=60)
        for f in ("queryIdx", "trainIdx", "distance"):
            assert np.array_equal(m[f], om[f]), f
    finally:
        fe.close()


def test_fast_score_ties_beyond_the_candidate_list_are_cut_deterministically():
    """A lattice of identical bright dots makes every corner score the same: the ties at the retainBest(2q) cut exceed the shared candidate list.
    That used to fail the frame with SFE_ERR_CAPACITY after an order-dependent truncation; now the ties with the smallest (y, x) are
    kept (documented deviation: OpenCV would carry all of them into the Harris stage), the frame succeeds and is reproducible."""
    from srba_stereo_slam_b200 import frontend
    W, H = 1241, 376
    yy, xx = np.mgrid[0:H, 0:W]
    img = np.where(((yy % 8) == 4) & ((xx % 8) == 4), 220, 20).astype(np.uint8)        # single bright pixels on an 8-px lattice: ~7000 corners of score 199
    fe = frontend.StereoFrontEnd(width=W, height=H, nlevels=1, nfeatures=300, max_pairs=1)
    try:
        outs = []
        for _ in range(3):
            fe.process_stereo(img, img)                          # must not raise SfeError(CAPACITY)
            kl, kr, dl, dr, m = fe.get_values(0)
            outs.append((kl.copy(), dl.copy(), m.copy()))
        n_raw = len(fe.get_raw_corners(0, 0)[0])
        assert n_raw > 2 * 300 + 1024                            # far more equal-score corners than the list holds
        assert len(outs[0][0]) >= 300
        for o in outs[1:]:
            assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1]) and np.array_equal(o[2], outs[0][2])
        assert np.array_equal(outs[0][0], fe.get_values(0)[1])    # left == right image -> identical keypoints
    finally:
        fe.close()


def test_match_bf_batch_device_initialises_its_keys_and_empty_train_sets():
    """sfe_match_bf_batch_device resets the packed keys itself (the header promises 0xFFFFFFFF for an empty train set; the kernel
    merges its train splits with atomicMin): a poisoned output buffer must not leak through."""
    import ctypes as C
    import torch
    from srba_stereo_slam_b200 import frontend
    from oracle import oracle as orc
    rng = np.random.default_rng(2)
    nq = [40, 17, 33]; nt = [50, 0, 21]
    Q = rng.integers(0, 256, (sum(nq), 32), dtype=np.uint8); T = rng.integers(0, 256, (sum(nt), 32), dtype=np.uint8)
    fe = frontend.StereoFrontEnd(width=320, height=240, nlevels=1, nfeatures=100, max_pairs=1)
    try:
        dev = torch.device("cuda", 0)
        dq = torch.from_numpy(Q).to(dev); dt = torch.from_numpy(T).to(dev)
        qo = torch.tensor(np.concatenate([[0], np.cumsum(nq)]), dtype=torch.int32, device=dev)
        to = torch.tensor(np.concatenate([[0], np.cumsum(nt)]), dtype=torch.int32, device=dev)
        best = torch.full((sum(nq),), 12345, dtype=torch.int32, device=dev)          # poison
        torch.cuda.synchronize()
        vp = C.c_void_p
        fe._chk(fe.lib.sfe_match_bf_batch_device(fe.h, vp(dq.data_ptr()), vp(qo.data_ptr()), vp(dt.data_ptr()), vp(to.data_ptr()), 3, max(nq), vp(best.data_ptr())))
        fe.synchronize()
        b = best.cpu().numpy().view(np.uint32)
        o0 = orc.match_bf(Q[:40], T[:50]); o2 = orc.match_bf(Q[57:], T[50:])
        assert np.array_equal(b[:40], (o0["distance"].astype(np.uint32) << 16) | o0["trainIdx"].astype(np.uint32))
        assert np.all(b[40:57] == 0xFFFFFFFF)
        assert np.array_equal(b[57:], (o2["distance"].astype(np.uint32) << 16) | o2["trainIdx"].astype(np.uint32))
    finally:
        fe.close()
