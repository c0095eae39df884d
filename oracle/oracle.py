"""ctypes front door to the CPU oracle (oracle/frontend_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importers allowed: tests/, __graft_entry__.smoke(), bench.py's
cpu_baseline / --impl reference legs.  The product (srba-stereo-slam_b200/) never imports this.

Every function mirrors one orc_* symbol; see the C file for the reference file:line each follows.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfrontend_oracle.so")

KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                           ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])   # cv::KeyPoint, 28 B
DMATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"),
                         ("distance", "<f4")])                                            # cv::DMatch, 16 B
assert KEYPOINT_DTYPE.itemsize == 28 and DMATCH_DTYPE.itemsize == 16


def build(force=False):
    src = os.path.join(_HERE, "frontend_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_level_scale.restype = C.c_float
        _lib.orc_level_scale.argtypes = [C.c_int, C.c_float]
        _lib.orc_harris.restype = C.c_float
        _lib.orc_ic_angle.restype = C.c_float
        _lib.orc_fast_atan2.restype = C.c_float
        _lib.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        _lib.orc_bow_l1_raw.restype = C.c_double
        _lib.orc_brief_pattern.restype = C.POINTER(C.c_int8)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OrbParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("nlevels", C.c_int32), ("fast_threshold", C.c_int32),
                ("edge_threshold", C.c_int32), ("scale_factor", C.c_float), ("blur_mode", C.c_int32), ("pyramid_mode", C.c_int32)]


def level_sizes(W, H, nlevels, scale=1.2, cv4=False):
    lw = np.zeros(nlevels, np.int32); lh = np.zeros(nlevels, np.int32)
    (lib().orc_level_sizes_cv4 if cv4 else lib().orc_level_sizes)(W, H, nlevels, C.c_float(scale), _p(lw), _p(lh))
    return lw, lh


def level_quotas(nfeatures, nlevels, scale=1.2):
    q = np.zeros(nlevels, np.int32)
    lib().orc_level_quotas(nfeatures, nlevels, C.c_float(scale), _p(q))
    return q


def level_scale(level, scale=1.2):
    return float(lib().orc_level_scale(level, C.c_float(scale)))


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), dw, dh, dw)
    return dst


def resize_linear_exact(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orc_resize_linear_exact_u8(_p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), dw, dh, dw)
    return dst


def fast_score_map(img, th):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros(img.shape, np.uint8)
    lib().orc_fast_score_map(_p(img), img.shape[1], img.shape[0], img.shape[1], th, _p(out))
    return out


def fast9_nms(img, th):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    n = lib().orc_fast9_nms(_p(img), w, h, w, th, None, None, None, 0)
    xs = np.zeros(n, np.int32); ys = np.zeros(n, np.int32); sc = np.zeros(n, np.int32)
    lib().orc_fast9_nms(_p(img), w, h, w, th, _p(xs), _p(ys), _p(sc), n)
    return xs, ys, sc


def harris(img, x, y):
    img = np.ascontiguousarray(img, np.uint8)
    return float(lib().orc_harris(_p(img), img.shape[1], int(x), int(y)))


def ic_angle(img, x, y):
    img = np.ascontiguousarray(img, np.uint8)
    return float(lib().orc_ic_angle(_p(img), img.shape[1], int(x), int(y)))


def fast_atan2(y, x):
    return float(lib().orc_fast_atan2(C.c_float(y), C.c_float(x)))


def umax():
    u = np.zeros(16, np.int32)
    lib().orc_umax(_p(u))
    return u


def gauss7_taps():
    t = np.zeros(7, np.int32)
    lib().orc_gauss7_taps(_p(t))
    return t


def gauss7_blur(img, blur_mode=0):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros_like(img)
    (lib().orc_gauss7_blur_f32path_u8 if blur_mode == 1 else lib().orc_gauss7_blur_u8)(_p(img), img.shape[1], img.shape[0], img.shape[1], _p(out), img.shape[1])
    return out


def brief(blurred, x, y, angle_deg):
    blurred = np.ascontiguousarray(blurred, np.uint8)
    d = np.zeros(32, np.uint8)
    lib().orc_brief(_p(blurred), blurred.shape[1], int(x), int(y), C.c_float(angle_deg), _p(d))
    return d


def brief_pattern():
    return np.ctypeslib.as_array(lib().orc_brief_pattern(), shape=(256, 4)).copy()


def orb_extract(img, nfeatures=1000, nlevels=8, fast_threshold=20, edge_threshold=31, scale=1.2,
                want_levels=False, blur_mode=0, pyramid_mode=0):
    """Whole-image ORB in canonical order.  Returns (keypoints[KEYPOINT_DTYPE], desc[n,32], extras)."""
    img = np.ascontiguousarray(img, np.uint8)
    H, W = img.shape
    prm = OrbParams(nfeatures, nlevels, fast_threshold, edge_threshold, scale, blur_mode, pyramid_mode)
    lw, lh = level_sizes(W, H, nlevels, scale, cv4=(pyramid_mode == 1))
    P = int((lw.astype(np.int64) * lh).sum())
    cap = int(4 * nfeatures + 4096)
    while True:
        kps = np.zeros(cap, KEYPOINT_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        counts = np.zeros(nlevels, np.int32)
        pyr = np.zeros(P, np.uint8) if want_levels else None
        blur = np.zeros(P, np.uint8) if want_levels else None
        n = lib().orc_orb_extract(_p(img), W, H, W, C.byref(prm), _p(kps), _p(desc), cap,
                                  _p(pyr) if want_levels else None, _p(blur) if want_levels else None,
                                  _p(counts))
        if n <= cap:
            break
        cap = n
    extras = {"level_counts": counts, "lw": lw, "lh": lh}
    if want_levels:
        offs = np.concatenate([[0], np.cumsum(lw.astype(np.int64) * lh)])
        extras["pyramid"] = [pyr[offs[l]:offs[l + 1]].reshape(lh[l], lw[l]) for l in range(nlevels)]
        extras["blurred"] = [blur[offs[l]:offs[l + 1]].reshape(lh[l], lw[l]) for l in range(nlevels)]
    return kps[:n].copy(), desc[:n].copy(), extras


def match_bf(Q, T):
    Q = np.ascontiguousarray(Q, np.uint8).reshape(-1, 32); T = np.ascontiguousarray(T, np.uint8).reshape(-1, 32)
    out = np.zeros(len(Q), DMATCH_DTYPE)
    n = lib().orc_match_bf(_p(Q), len(Q), _p(T), len(T), _p(out))
    return out[:n]


def match_rowband(kl, dl, kr, dr, max_y_diff=2.0, min_disp=0.0, max_disp=1e9, max_dist=60, robust_1to1=False):
    kl = np.ascontiguousarray(kl, KEYPOINT_DTYPE); kr = np.ascontiguousarray(kr, KEYPOINT_DTYPE)
    dl = np.ascontiguousarray(dl, np.uint8); dr = np.ascontiguousarray(dr, np.uint8)
    out = np.zeros(len(kl), DMATCH_DTYPE)
    n = lib().orc_match_rowband(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr),
                                C.c_float(max_y_diff), C.c_float(min_disp), C.c_float(max_disp),
                                int(max_dist), int(bool(robust_1to1)), _p(out))
    return out[:n]


def bow_transform(voc, desc):
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    n = len(desc)
    word = np.zeros(n, np.int32); weight = np.zeros(n, np.float64)
    lib().orc_bow_transform(_p(voc["child_ofs"]), _p(voc["child_idx"]), _p(voc["node_desc"]),
                            _p(voc["node_weight"]), _p(voc["node_word"]), _p(desc), n, _p(word), _p(weight))
    return word, weight


def bow_vector(voc, word, weight):
    n = len(word)
    ow = np.zeros(max(n, 1), np.int32); ov = np.zeros(max(n, 1), np.float64)
    word = np.ascontiguousarray(word, np.int32); weight = np.ascontiguousarray(weight, np.float64)
    m = lib().orc_bow_vector(_p(word), _p(weight), n, int(voc["nwords"]), _p(ow), _p(ov))
    return ow[:m].copy(), ov[:m].copy()


def bow_l1_raw(qw, qv, dw, dv):
    common = C.c_int(0)
    qw = np.ascontiguousarray(qw, np.int32); dw = np.ascontiguousarray(dw, np.int32)
    qv = np.ascontiguousarray(qv, np.float64); dv = np.ascontiguousarray(dv, np.float64)
    s = lib().orc_bow_l1_raw(_p(qw), _p(qv), len(qw), _p(dw), _p(dv), len(dw), C.byref(common))
    return float(s), bool(common.value)


def bow_query(voc, db_vectors, desc, max_results=4):
    """DBoW2 TemplatedDatabase::query with L1 scoring: [(entry_id, score)] best first; ties -> lower id."""
    w, wt = bow_transform(voc, desc)
    qw, qv = bow_vector(voc, w, wt)
    res = []
    for eid, (dw, dv) in enumerate(db_vectors):
        s, common = bow_l1_raw(qw, qv, dw, dv)
        if common:
            res.append((s, eid))
    res.sort()
    if max_results > 0:
        res = res[:max_results]
    return [(eid, -s / 2.0) for s, eid in res]
