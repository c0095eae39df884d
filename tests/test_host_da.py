"""CPU: the host-side data-association steps of SURVEY.md 8(f) (direction filter, ORB distance + uniqueness filter with
the reference's counter bug, match-id assignment, projectMatchTo3D) in libsrba_b200_host.so against the line-by-line
oracle restatement (oracle/da_oracle.py)."""
import ctypes as C
import os
import numpy as np
import pytest

from oracle import da_oracle as dao

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SO = os.path.join(ROOT, "srba-stereo-slam_b200", "host", "libsrba_b200_host.so")


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(HOST_SO):
        pytest.skip("host adapters not built (run __graft_entry__.build())")
    return C.CDLL(HOST_SO)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_direction_filter(host):
    rng = np.random.default_rng(0)
    for trial in range(20):
        n = int(rng.integers(5, 300))
        txy = rng.uniform(30, 1200, (n, 2)).astype(np.float32); txy[:, 1] = rng.uniform(30, 340, n)
        flow = rng.normal([6.0, 1.0], 2.0, (n, 2)).astype(np.float32)
        oxy = (txy + flow).astype(np.float32)
        out = rng.random(n) < 0.25                                    # gross outliers
        oxy[out] = rng.uniform(30, 1200, (int(out.sum()), 2)).astype(np.float32) * np.float32([1, 0.28])
        if trial % 5 == 0:
            oxy[0, 0] = txy[0, 0]                                     # vertical displacement: atanf(+-inf)
        st0 = np.where(rng.random(n) < 0.1, dao.REJ_ORB, dao.TRACKED).astype(np.int32)
        st = st0.copy()
        host.srba_da_filter_direction(_p(txy), _p(oxy), n, C.c_float(370.0), _p(st))
        exp = dao.filter_direction(txy, oxy, 370.0, st0)
        assert st.tolist() == exp
        assert (st == dao.REJ_SLOPE).sum() > 0 or n < 10


def test_orb_distance_filter_reproduces_counter_bug(host):
    host.srba_da_filter_orb_distance.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p]
    rng = np.random.default_rng(1)
    for trial in range(30):
        n = int(rng.integers(3, 400)); n_other = int(rng.integers(3, 400))
        q = np.arange(n, dtype=np.int32)
        t = rng.integers(0, n_other, n).astype(np.int32)              # many-to-one on purpose
        d = rng.integers(0, 120, n).astype(np.float32)
        st0 = np.where(rng.random(n) < 0.15, dao.REJ_SLOPE, dao.TRACKED).astype(np.int32)
        st = st0.copy()
        host.srba_da_filter_orb_distance(_p(q), _p(t), _p(d), n, n_other, 60.0, _p(st))
        assert st.tolist() == dao.filter_orb_distance(q, t, d, n_other, 60.0, st0)
    # the bug, pinned on a hand-made case: two queries hit train 0; the slot holds the COUNTER (1.0), so a later match at
    # distance 0 replaces the earlier one, a later match at distance 5 (> 1.0) does not
    q = np.arange(3, dtype=np.int32); t = np.zeros(3, np.int32)
    st = np.zeros(3, np.int32); d = np.array([50, 5, 0], np.float32)
    host.srba_da_filter_orb_distance(_p(q), _p(t), _p(d), 3, 1, 60.0, _p(st))
    assert st.tolist() == [dao.NON_TRACKED, dao.NON_TRACKED, dao.TRACKED]
    st = np.zeros(2, np.int32); d = np.array([50, 5], np.float32)      # a correct implementation would keep query 1
    host.srba_da_filter_orb_distance(_p(q), _p(t), _p(d), 2, 1, 60.0, _p(st))
    assert st.tolist() == [dao.TRACKED, dao.NON_TRACKED]


def test_match_id_assignment(host):
    rng = np.random.default_rng(2)
    for trial in range(20):
        n_kfs = int(rng.integers(2, 8)); nm = int(rng.integers(1, 200)); nsim = int(rng.integers(1, min(5, n_kfs) + 1))
        kf_ids = [rng.permutation(5000)[:int(rng.integers(5, 300))].tolist() for _ in range(n_kfs)]
        da_kf = rng.permutation(n_kfs)[:nsim].astype(np.int32)
        first = np.full((nsim, nm), -1, np.int64)
        for k in range(nsim):
            hit = rng.random(nm) < 0.4
            first[k, hit] = rng.integers(0, len(kf_ids[da_kf[k]]), int(hit.sum()))
        srt = rng.permutation(nsim).astype(np.int32)
        flat = np.array([i for ids in kf_ids for i in ids], np.int64)
        ofs = np.concatenate([[0], np.cumsum([len(i) for i in kf_ids])]).astype(np.int32)
        last = C.c_longlong(10_000); nt = C.c_longlong(); nn = C.c_longlong()
        out = np.zeros(nm, np.int64)
        host.srba_da_assign_ids(nm, nsim, _p(first), _p(da_kf), _p(srt), _p(flat), _p(ofs), n_kfs, C.byref(last), _p(out), C.byref(nt), C.byref(nn))
        exp, elast, et, en = dao.assign_match_ids(nm, [(int(da_kf[k]), first[k].tolist()) for k in range(nsim)], srt.tolist(), kf_ids, 10_000)
        assert out.tolist() == exp and last.value == elast and (nt.value, nn.value) == (et, en)
        assert len(set(out.tolist())) == nm                            # ids are unique within a keyframe


def test_project_match_to_3d(host):
    cam = np.array([601.8873, 183.1104, 707.0912, 601.8873, 707.0912, 0.54])       # kitti.ini:140-155
    xyz = np.zeros(3)
    host.srba_project_match_to_3d.argtypes = [C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    for ul, vl, ur in ((700.0, 200.0, 680.0), (100.5, 50.25, 99.0), (601.8873, 183.1104, 560.0)):
        host.srba_project_match_to_3d(ul, vl, ur, _p(cam), _p(xyz))
        assert xyz.tolist() == list(dao.project_match_to_3d(ul, vl, ur, *cam))
    host.srba_project_match_to_3d(700.0, 200.0, 680.0, _p(cam), _p(xyz))
    assert abs(xyz[2] - 0.54 * 707.0912 / 20.0) < 1e-9                 # Z = B f / disparity


def test_keyframe_stream_format_is_byte_compatible(host, tmp_path):
    """EST.cpp:2517-2616: size_t n; {x, y, response, size, angle, octave, class_id}; int rows, cols, type; bytes ...
    then matches as {size_t id; int queryIdx, trainIdx; float distance; int imgIdx}.  Written by the C++ code, parsed
    here byte by byte with the reference's layout, then loaded back by the C++ code."""
    import struct
    from srba_stereo_slam_b200 import frontend
    rng = np.random.default_rng(5)
    nl, nr, nm = 7, 5, 4

    def kps(n):
        k = np.zeros(n, frontend.KEYPOINT_DTYPE)
        for f in ("x", "y", "size", "angle", "response"):
            k[f] = rng.uniform(0, 400, n).astype(np.float32)
        k["octave"] = rng.integers(0, 8, n); k["class_id"] = -1
        return k
    kl, kr = kps(nl), kps(nr)
    dl = rng.integers(0, 256, (nl, 32), dtype=np.uint8); dr = rng.integers(0, 256, (nr, 32), dtype=np.uint8)
    m = np.zeros(nm, frontend.DMATCH_DTYPE); m["queryIdx"] = [0, 2, 3, 6]; m["trainIdx"] = [1, 0, 4, 2]; m["distance"] = [3, 17, 40, 59]
    ids = np.array([100, 7, 2**40 + 5, 3], np.uint64)
    path = str(tmp_path / "kf.bin").encode()
    assert host.srba_kf_dump_file(path, _p(kl), nl, _p(kr), nr, _p(dl), _p(dr), _p(m), nm, _p(ids)) == 0
    raw = open(path, "rb").read(); off = 0
    for k, d in ((kl, dl), (kr, dr)):
        (n,) = struct.unpack_from("<Q", raw, off); off += 8
        assert n == len(k)
        for i in range(n):
            x, y, resp, size, ang, octv, cid = struct.unpack_from("<5f2i", raw, off); off += 28
            assert (x, y, resp, size, ang, octv, cid) == (k["x"][i], k["y"][i], k["response"][i], k["size"][i], k["angle"][i], k["octave"][i], k["class_id"][i])
        rows, cols, typ = struct.unpack_from("<3i", raw, off); off += 12
        assert (rows, cols, typ) == (len(k), 32, 0)                         # CV_8U == 0
        assert raw[off:off + rows * cols] == d.tobytes(); off += rows * cols
    (n,) = struct.unpack_from("<Q", raw, off); off += 8
    assert n == nm
    for i in range(nm):
        mid, q, t, dist, img = struct.unpack_from("<Q2ifi", raw, off); off += 24
        assert (mid, q, t, dist, img) == (int(ids[i]), m["queryIdx"][i], m["trainIdx"][i], m["distance"][i], m["imgIdx"][i])
    assert off == len(raw)
    kl2, kr2 = np.zeros(16, frontend.KEYPOINT_DTYPE), np.zeros(16, frontend.KEYPOINT_DTYPE)
    dl2, dr2 = np.zeros((16, 32), np.uint8), np.zeros((16, 32), np.uint8)
    m2 = np.zeros(16, frontend.DMATCH_DTYPE); ids2 = np.zeros(16, np.uint64)
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    assert host.srba_kf_load_file(path, _p(kl2), C.byref(a), _p(kr2), C.byref(b), _p(dl2), _p(dr2), _p(m2), C.byref(c), _p(ids2), 16) == 0
    assert (a.value, b.value, c.value) == (nl, nr, nm)
    assert np.array_equal(kl2[:nl], kl) and np.array_equal(kr2[:nr], kr) and np.array_equal(dl2[:nl], dl) and np.array_equal(dr2[:nr], dr)
    assert np.array_equal(m2[:nm], m) and np.array_equal(ids2[:nm], ids)
