"""GPU (-m gpu): sequence-level parity (BASELINE.json configs[0]: demo/config_imgdir_kitti_srba.ini on a 200-frame synthetic
1241x376 stereo sequence).

SRBA/MRPT are not available, so the final keyframe poses cannot be compared; what CAN be compared, and is what those poses are
a function of, is the input of rba.define_new_keyframe (CSRBAStereoSLAMEstimator.cpp:135-166, 666-782): one observation list
{feat_id, left_px, right_px, initial 3-D point} per keyframe.  The C++ adapters + CUDA library (host/slam_sequence over
CSRBAStereoSLAMEstimator, rso::CStereoOdometryEstimator, CStereoSLAMKF, CBoWManager) run kitti.ini's values over the frames;
oracle/slam_oracle.py runs the oracle chain over the same frames; every keyframe decision and every observation must agree.
A second test runs the REFERENCE'S OWN front-end loop (oracle/_ref, compiled from /root/reference by oracle/ref_build.py:
initialize, first frame, adaptive-threshold loop, pose accumulation, candidate keyframe + BoW query) on the GPU through the
drop-in headers and holds it to the same oracle.
"""
import ctypes as C
import os
import struct
import subprocess
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "srba-stereo-slam_b200", "host")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_da.so")

KITTI_INI = """
[RECTIFY]
nOctaves = 1
[DETECT]
min_distance = 5
non_maximal_suppression = true
non_max_supp_method = 1
[MATCH]
max_y_diff = 2
enable_robust_1to1_match = false
rectified_images = true
orb_max_distance = 60
[IF-MATCH]
filter_fund_matrix = false
[LEAST_SQUARES]
initial_max_iters = 30
max_iters = 30
max_incr_cost = 3
residual_threshold = 15
min_mod_out_vector = 0.001
bad_tracking_th = 5
use_robust_kernel = true
kernel_param = 2
use_previous_pose_as_initial = true
[GENERAL]
camera_pose_on_robot = [0,0,0,0,0,0]
[APP_OPTIONS]
out_dir = {out_dir}
debug = {debug}
show3D = false
verbose_level = 0
max_num_kfs = 500
[SRBA_GENERAL]
voc_filename = {voc}
srba_max_tree_depth = 4
srba_max_optimize_depth = 5
srba_submap_size = 10
[SRBA_DETECT]
n_feats = 500
orb_adaptive_fast_th = true
detect_fast_th = 20
adaptive_th_min_matches = 100
[SRBA_DATA_ASSOCIATION]
da_stage2_method = 3
residual_th = 30
max_orb_distance_da = 60
max_y_diff_epipolar = 2.0
ransac_fit_prob = 0.95
da_filter_by_direction = true
[SRBA_KF_CREATION]
max_rotation = 25
max_translation = 2.0
updated_matches_th = 60
up_matches_th_plus = 25
lc_distance = 10
vo_id_tracking_th = 40
use_initial_pose = true
[CAMERA_LEFT]
resolution = [{W} {H}]
cx = {cx!r}
cy = {cy!r}
fx = {fx!r}
fy = {fy!r}
dist = [0 0 0 0 0]
[CAMERA_RIGHT]
resolution = [{W} {H}]
cx = {cx!r}
cy = {cy!r}
fx = {fx!r}
fy = {fy!r}
dist = [0 0 0 0 0]
[CAMERA_LEFT2RIGHT_POSE]
pose_quaternion = [{b!r} 0 0 1 0 0 0]
"""
OPT = dict(n_feats=500, n_levels=1, detect_fast_th=20, orb_adaptive_fast_th=True, adaptive_th_min_matches=100, da_stage2_method=3, residual_th=30,
           max_orb_distance_da=60, max_y_diff_epipolar=2.0, ransac_fit_prob=0.95, query_score_th=0.04, da_filter_by_direction=True,
           da_filter_by_orb_distance=True, da_filter_by_fund_matrix=True, da_filter_by_pose_change=True, max_rotation=25, max_translation=2.0,
           updated_matches_th=60, up_matches_th_plus=25, lc_distance=10, vo_id_tracking_th=40, use_initial_pose=True, srba_submap_size=10,
           srba_max_tree_depth=4, srba_max_optimize_depth=5)                     # = the [SRBA_*] blocks of demo/config_imgdir_kitti_srba.ini
LS = dict(initial_max_iters=30, max_iters=30, max_incr_cost=3, residual_threshold=15, min_mod_out_vector=0.001, bad_tracking_th=5,
          use_robust_kernel=True, kernel_param=2.0)


def render_sequence(n, W, H, radius, seed, tmp_path):
    """Renders in a fresh child process (which forks its own worker pool): forking workers from the multi-threaded pytest process
    itself risks a deadlock."""
    import sys
    out = tmp_path / "render.bin"
    synth_py = os.path.join(ROOT, "srba-stereo-slam_b200", "synth.py")
    subprocess.check_call([sys.executable, synth_py, str(n), str(W), str(H), repr(float(radius)), str(seed), str(out)], timeout=1500)
    fr = np.fromfile(out, np.uint8).reshape(n, 2, H, W)
    os.remove(out)
    return fr


def write_voc_yml(voc, path):
    from test_gpu_host_adapters import _write_voc_yml
    _write_voc_yml(voc, path)


def parse_sequence_dump(path):
    data = open(path, "rb").read(); off = 0

    def rd(fmt):
        nonlocal off
        v = struct.unpack_from("=" + fmt, data, off); off += struct.calcsize("=" + fmt)
        return v
    nkf, = rd("i"); lists = []; frames = []
    for _ in range(nkf):
        f, n = rd("ii"); frames.append(f)
        rec = np.frombuffer(data, np.dtype([("id", "<i8"), ("px", "<f4", 4), ("xyz", "<f8", 3)]), n, off); off += n * 48
        lists.append(rec)
    nck, = rd("i"); checks = []
    for _ in range(nck):
        frame, cand, flags, lc_id, ntr, nnew = rd("6i")
        nq, = rd("i"); q = [rd("id") for _ in range(nq)]
        ns, = rd("i"); sim = list(rd("%di" % ns)) if ns else []
        nt, = rd("i"); tr = [rd("ii") for _ in range(nt)]
        checks.append(dict(frame=frame, cand=cand, flags=flags, lc_id=lc_id, n_tracked_ids=ntr, n_new_ids=nnew, query=q, similar=sim, tracked=tr))
    npose, = rd("i")
    poses = np.frombuffer(data, np.float64, 6 * npose, off).reshape(npose, 6); off += 48 * npose
    assert off == len(data)
    return lists, frames, checks, poses


def run_sequence_case(tmp_path, voc, W, H, NF, radius, seed):
    from srba_stereo_slam_b200 import synth
    from oracle import slam_oracle as so
    sc = synth.CorridorScene(seed=seed, width=W, height=H, loop_radius=radius)
    frames = render_sequence(NF, W, H, radius, seed, tmp_path)
    fb = tmp_path / "frames.bin"; frames.tofile(fb)
    vy = tmp_path / "voc.yml"; write_voc_yml(voc, vy)
    ini = tmp_path / "kitti_synth.ini"
    ini.write_text(KITTI_INI.format(out_dir=str(tmp_path / "out"), debug="false", voc=str(vy), W=W, H=H, cx=float(sc.cx), cy=float(sc.cy), fx=float(sc.fx), fy=float(sc.fy), b=float(sc.b)))
    ob = tmp_path / "seq.bin"
    r = subprocess.run([os.path.join(HOST, "slam_sequence"), str(ini), str(fb), str(W), str(H), str(NF), str(ob)], capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stdout + r.stderr
    lists, kf_frames, checks, poses = parse_sequence_dump(ob)
    cam = (float(sc.cx), float(sc.cy), float(sc.fx), float(sc.cx), float(sc.fx), float(sc.b))
    o = so.SlamOracle(voc, OPT, cam, H, LS).run(frames, max_num_kfs=500)
    return sc, frames, lists, kf_frames, checks, poses, o


def compare_with_oracle(lists, kf_frames, checks, poses, o):
    # ---- every keyframe-check decision -------------------------------------------------------------------------------------
    assert len(checks) == len(o.checks), (len(checks), len(o.checks))
    for c, oc in zip(checks, o.checks):
        fl = oc["flags"]
        oflags = int(fl[0]) | int(fl[1]) << 1 | int(fl[2]) << 2 | int(oc["potential_lc"]) << 3 | int(oc["confirmed_lc"]) << 4 | int(oc["inserted"]) << 5
        assert (c["frame"], c["flags"]) == (oc["frame"], oflags), (c, oc)
        assert [q[0] for q in c["query"]] == [q[0] for q in oc["query"]] and [q[1] for q in c["query"]] == [q[1] for q in oc["query"]], c["frame"]   # FP64 scores, bit for bit
        assert c["similar"] == [int(s) for s in oc["similar"]] and c["lc_id"] == oc["lc_id"], c["frame"]
        assert [tuple(t) for t in c["tracked"]] == [tuple(t) for t in oc["tracked"]], c["frame"]
        if oc["inserted"]:
            assert (c["n_tracked_ids"], c["n_new_ids"]) == (oc["n_tracked_ids"], oc["n_new_ids"]), c["frame"]
    # ---- every observation list (the argument of rba.define_new_keyframe) ---------------------------------------------------------
    assert kf_frames == o.kf_frames
    assert len(lists) == len(o.obs_lists)
    nobs = 0
    for k, (a, b) in enumerate(zip(lists, o.obs_lists)):
        assert len(a) == len(b), k
        bb = np.array(b, np.float64).reshape(-1, 8)
        assert np.array_equal(a["id"], bb[:, 0].astype(np.int64)), k                           # landmark ids
        assert np.array_equal(a["px"].astype(np.float64), bb[:, 1:5]), k                       # left / right pixel coordinates
        assert np.allclose(a["xyz"], bb[:, 5:8], rtol=1e-12, atol=1e-12), k                     # initial landmark positions
        nobs += len(a)
    # ---- the odometry poses the decisions were made from (north_star tolerance: 1e-3 m / 1e-3 rad; observed: ~1e-12) ----------------
    op = np.array([[p[3], p[4], p[5], p[0], p[1], p[2]] for p in o.frame_poses])                # oracle order (w, t) -> dump order (x y z wx wy wz)
    assert poses.shape == op.shape
    assert np.abs(poses - op).max() < 1e-6
    return nobs


def test_sequence_200_frames_config0(tmp_path, voc):
    """configs[0]: kitti.ini as shipped (500 ORB, n_levels = 1, FAST 20 adaptive, all four DA filters, stage-2 method 3) on 200 synthetic
    1241x376 frames; the ring closes after ~157 frames, so the run revisits its start (loop-closure candidates in the BoW query)."""
    W, H, NF = 1241, 376, int(os.environ.get("SRBA_SEQ_FRAMES", "200"))
    sc, frames, lists, kf_frames, checks, poses, o = run_sequence_case(tmp_path, voc, W, H, NF, 20.0, 5)
    nobs = compare_with_oracle(lists, kf_frames, checks, poses, o)
    assert len(lists) >= 20 and nobs > 4000 and len(checks) >= len(lists) - 1
    tracked_total = sum(c["n_tracked_ids"] for c in checks)
    assert tracked_total > 500                                   # the data association really links keyframes
    # ground truth: the dead-reckoned camera positions stay near the true ring (sanity of the whole chain, not a parity claim)
    pos = poses[:, :3]
    step = np.linalg.norm(np.diff(pos, axis=0), axis=1)
    assert 0.5 < np.median(step) < 1.1                           # 0.8 m per frame
    if NF >= 170:
        revisit = [c for c in checks if c["frame"] >= 160 and any(q[0] < 8 for q in c["query"])]
        assert revisit, "no BoW hit on the keyframes of the first lap after the ring closed"


def test_reference_loop_runs_on_the_gpu_through_the_drop_in_headers(tmp_path, voc):
    """oracle/_ref/libref_da.so = the reference's own lines (initialize :1102-1144,:1163-1167; first frame :111-118; adaptive loop + pose
    accumulation :258-330; candidate keyframe + queryDB :413-427), compiled unchanged against host/*.h, driving the CUDA library."""
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libref_da.so did not travel with the snapshot")
    from srba_stereo_slam_b200 import synth
    from oracle import slam_oracle as so, oracle as orc, geometry_oracle as geo
    W, H, NF = 640, 192, 14
    sc = synth.CorridorScene(seed=6, width=W, height=H, loop_radius=40.0)
    frames = sc.sequence(NF)
    frames[5] = (frames[5].astype(np.float32) * 0.3 + 90).astype(np.uint8)      # low contrast: the adaptive loop must repeat
    vy = tmp_path / "voc.yml"; write_voc_yml(voc, vy)
    ini = tmp_path / "k.ini"
    ini.write_text(KITTI_INI.format(out_dir=str(tmp_path / "out"), debug="false", voc=str(vy), W=W, H=H, cx=float(sc.cx), cy=float(sc.cy), fx=float(sc.fx), fy=float(sc.fy), b=float(sc.b)))
    L = C.CDLL(REF_SO)
    L.ref_vo_open.restype = C.c_void_p; L.ref_vo_open.argtypes = [C.c_char_p]
    L.ref_vo_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
    L.ref_vo_close.argtypes = [C.c_void_p]
    h = L.ref_vo_open(str(ini).encode())
    assert h
    cam = (float(sc.cx), float(sc.cy), float(sc.fx), float(sc.cx), float(sc.fx), float(sc.b))
    chain = so.SlamOracle(voc, OPT, cam, H, LS)
    repeats = 0
    try:
        for k in range(NF):
            oi = np.zeros(9, np.int32); pose = np.zeros(6); ids = np.zeros(4096, np.int64); qi = np.zeros(4, np.int32); qs = np.zeros(4)
            l = np.ascontiguousarray(frames[k, 0]); r = np.ascontiguousarray(frames[k, 1])
            insert = int(k % 3 == 0)
            nm = L.ref_vo_frame(h, l.ctypes.data, r.ctypes.data, W, H, insert, oi.ctypes.data, pose.ctypes.data, ids.ctypes.data, qi.ctypes.data, qs.ctypes.data)
            assert nm >= 0
            # ---- the oracle, same steps
            if k == 0:
                chain.vo.process(l, r); chain.vo.set_this_frame_as_kf()
                kf = chain.vo.get_values(); chain._bow_insert(kf); valid = True; res = dict(tracked=0, tracked_kf=0); q = []
            else:
                f0 = chain.vo.fast_th
                res = chain._adaptive(l, r); valid = res["valid"]
                repeats += int(chain.vo.fast_th < f0 - 5)
                q = []
                if valid:
                    chain.cur_pose = geo.rv_compose(chain.cur_pose, res["out_pose"])
                    kf = chain.vo.get_values()
                    q = orc.bow_query(voc, chain.db, kf["dl"], 4)
                    if insert:
                        chain._bow_insert(kf)
            cur = chain.vo.cur
            assert oi[0] == int(valid), k
            assert (oi[1], oi[2], oi[3]) == (len(cur["kl"]), len(cur["kr"]), len(cur["m"])), k
            assert np.array_equal(ids[:nm], np.array(cur["ids"], np.int64)), k
            assert (oi[6], oi[7]) == (chain.vo.fast_th, chain.vo.orb_th), k
            if k > 0:
                assert (oi[4], oi[5]) == (res["tracked"], res["tracked_kf"]), k
            if valid and k > 0:
                assert oi[8] == len(q) and [int(x) for x in qi[:oi[8]]] == [x[0] for x in q] and [float(x) for x in qs[:oi[8]]] == [x[1] for x in q], k
            ocp = chain.cur_pose
            assert np.allclose(pose, [ocp[3], ocp[4], ocp[5], ocp[0], ocp[1], ocp[2]], rtol=0, atol=1e-8), k
    finally:
        L.ref_vo_close(h)
    assert repeats >= 1


def test_debug_dump_formats(tmp_path, voc):
    """[APP_OPTIONS] debug = true: the ten text artefacts of SURVEY.md Appendix C with the reference's formats -- info_kf / info_feats_kf
    (CStereoSLAMKF.cpp:60-110), if_raw_match / if_match_after (EST.cpp:1455-1473, 1650-1689), da_dist (:602, :614), da_info (:749-764),
    da.txt (:649), ls.txt (:790), time_new_kf.txt (:939-942), out_kf_poses.txt (:977-986) -- cross-checked against the binary dump of the same run."""
    from srba_stereo_slam_b200 import synth
    W, H, NF = 640, 192, 24
    sc = synth.CorridorScene(seed=6, width=W, height=H, loop_radius=40.0)
    frames = sc.sequence(NF)
    fb = tmp_path / "frames.bin"; frames.tofile(fb)
    vy = tmp_path / "voc.yml"; write_voc_yml(voc, vy)
    out = tmp_path / "out"
    ini = tmp_path / "k.ini"
    ini.write_text(KITTI_INI.format(out_dir=str(out), debug="true", voc=str(vy), W=W, H=H, cx=float(sc.cx), cy=float(sc.cy), fx=float(sc.fx), fy=float(sc.fy), b=float(sc.b)))
    ob = tmp_path / "seq.bin"
    r = subprocess.run([os.path.join(HOST, "slam_sequence"), str(ini), str(fb), str(W), str(H), str(NF), str(ob)], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    lists, kf_frames, checks, poses = parse_sequence_dump(ob)
    nkf = len(lists)
    assert nkf >= 3
    for k in range(nkf):
        rows = [l.split() for l in open(out / ("info_kf%04d.txt" % k)).read().splitlines()]
        assert rows[0][0] == "%" and len(rows) - 1 == len(lists[k])
        for row, o in zip(rows[1:], lists[k]):
            assert int(row[0]) == k and int(row[1]) == int(o["id"]) and [float(v) for v in row[2:6]] == [round(float(v), 2) for v in o["px"]]
        feats = open(out / ("info_feats_kf%04d.txt" % k)).read().split()
        nl, nr = int(feats[0]), int(feats[1]); assert len(feats) == 2 + 2 * (nl + nr)
        if k > 0:
            da = [l.split() for l in open(out / ("da_info_%04d.txt" % k)).read().splitlines()]
            assert [int(r_[0]) for r_ in da] == [int(v) for v in lists[k]["id"]]
            assert [[float(v) for v in r_[1:]] for r_ in da] == [[round(float(v), 2) for v in o["px"]] for o in lists[k]]
            dd = open(out / ("da_dist_kf%04d.txt" % k)).read().splitlines()
            assert len(dd) == len(lists[k])                      # one line per stereo match: the tracked distance, or 0.00 for a new landmark
    ins = [c for c in checks if c["flags"] & 32]
    for c in ins:                                                # one raw / filtered match file per associated keyframe
        for kf, _ in c["tracked"]:
            raw = open(out / ("if_raw_match_kf%04d_with_kf%04d.txt" % (c["cand"], kf))).read().splitlines()
            assert raw[0].startswith("% OTHER_LX") and all(len(l.split()) == 5 for l in raw[1:])
    after = [f for f in os.listdir(out) if f.startswith("if_match_after_")]
    assert after and all(len(l.split()) == 6 for l in open(out / after[0]).read().splitlines()[1:])
    da_txt = [l.split() for l in open(tmp_path / "da.txt").read().splitlines()]
    assert sum(len(c["tracked"]) for c in ins) == len(da_txt) and all(len(r_) == 3 for r_ in da_txt)
    ls = [l.split() for l in open(tmp_path / "ls.txt").read().splitlines()]
    assert [int(r_[0]) for r_ in ls] == list(range(1, nkf))
    tk = [l.split() for l in open(tmp_path / "time_new_kf.txt").read().splitlines()]
    assert [int(r_[1]) for r_ in tk] == [len(l) for l in lists]
    kp = [l.split() for l in open(tmp_path / "out_kf_poses.txt").read().splitlines()]
    assert [int(r_[0]) for r_ in kp] == list(range(nkf)) and all(len(r_) == 7 for r_ in kp)
