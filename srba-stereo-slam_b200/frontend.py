"""ctypes binding of include/srba_b200.h (the drop-in boundary).

Mirrors the reference-facing stage contract one to one:
  StereoFrontEnd.process_stereo   <- rso::CStereoOdometryEstimator::processNewImagePair (EST.cpp:112,267)
  StereoFrontEnd.get_values       <- voEngine.getValues (CStereoSLAMKF.h:68-70)
  set/get_fast_threshold, set/get_orb_threshold <- EST.cpp:275-311
  StereoFrontEnd.match_bf         <- cv::BFMatcher(NORM_HAMMING,false).match (EST.cpp:1447-1452)
  BowDatabase.add / query         <- CBoWManager::insertIntoDB / queryDB (CBoWManager.h:71-88)
No compute happens here and there is no fallback: a missing library or GPU raises.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsrba_b200.so")

KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                           ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
DMATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
COUNTS_DTYPE = np.dtype([("n_left", "<i4"), ("n_right", "<i4"), ("n_matches", "<i4"), ("status", "<i4")])

SFE_OK, SFE_ERR_INVALID, SFE_ERR_CUDA, SFE_ERR_CAPACITY, SFE_ERR_STATE = range(5)
BLUR_CV24_FIXED, BLUR_FLOAT = 0, 1
PYRAMID_CV24_LINEAR, PYRAMID_CV4_EXACT = 0, 1


class SfeConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("nlevels", C.c_int32), ("nfeatures", C.c_int32),
                ("scale_factor", C.c_float), ("edge_threshold", C.c_int32), ("fast_threshold", C.c_int32),
                ("blur_mode", C.c_int32), ("max_pairs", C.c_int32), ("max_y_diff", C.c_float),
                ("min_disparity", C.c_float), ("max_disparity", C.c_float), ("orb_max_distance", C.c_int32),
                ("robust_1to1", C.c_int32), ("device", C.c_int32), ("use_graph", C.c_int32), ("pyramid_mode", C.c_int32)]


class SfeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("sfe error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load_library():
    """Load libsrba_b200.so; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libsrba_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                          "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i32, sz = C.c_void_p, C.c_int32, C.c_size_t
    L.sfe_version.restype = C.c_char_p
    L.sfe_last_error.restype = C.c_char_p; L.sfe_last_error.argtypes = [vp]
    L.sfe_default_config.argtypes = [C.POINTER(SfeConfig)]
    L.sfe_create.argtypes = [C.POINTER(SfeConfig), C.POINTER(vp)]
    L.sfe_destroy.argtypes = [vp]; L.sfe_destroy.restype = None
    L.sfe_stream.argtypes = [vp]; L.sfe_stream.restype = vp
    for n in ("sfe_set_fast_threshold", "sfe_set_orb_threshold"):
        getattr(L, n).argtypes = [vp, i32]
    for n in ("sfe_get_fast_threshold", "sfe_get_orb_threshold", "sfe_num_levels", "sfe_keypoint_capacity", "sfe_synchronize",
              "sfe_timer_start"):
        getattr(L, n).argtypes = [vp]
    L.sfe_level_size.argtypes = [vp, i32, vp, vp, vp, vp]
    L.sfe_process_batch_host.argtypes = [vp, vp, i32, sz, sz]
    L.sfe_process_stereo.argtypes = [vp, vp, vp, sz, i32]
    L.sfe_host_staging.argtypes = [vp, C.POINTER(sz)]; L.sfe_host_staging.restype = vp
    L.sfe_upload_batch.argtypes = [vp, vp, i32, sz, sz]
    L.sfe_process_batch_device.argtypes = [vp, i32]
    L.sfe_get_counts.argtypes = [vp, i32, vp]
    L.sfe_get_values.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
    L.sfe_download_batch.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.sfe_submit_batch_host.argtypes = [vp, vp, i32, sz, sz]
    L.sfe_wait_batch.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.sfe_get_level.argtypes = [vp, i32, i32, i32, vp]
    L.sfe_get_raw_corners.argtypes = [vp, i32, i32, vp, i32, vp]
    L.sfe_device_results.argtypes = [vp] + [C.POINTER(vp)] * 5
    L.sfe_match_bf.argtypes = [vp, vp, i32, vp, i32, vp, vp]
    L.sfe_match_bf_batch_device.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp]
    L.sfe_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.sfe_stream_wait.argtypes = [vp, vp]
    L.sfe_set_profiling.argtypes = [vp, i32]
    L.sfe_stage_name.argtypes = [i32]; L.sfe_stage_name.restype = C.c_char_p
    L.sfe_stage_times.argtypes = [vp, vp, vp]
    L.sfe_kernel_launches.argtypes = [vp]; L.sfe_kernel_launches.restype = C.c_int64
    L.sfe_bow_create.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, C.POINTER(vp)]
    L.sfe_bow_destroy.argtypes = [vp]; L.sfe_bow_destroy.restype = None
    L.sfe_bow_transform.argtypes = [vp, vp, i32, vp, vp]
    L.sfe_bow_vector.argtypes = [vp, vp, i32, vp, vp, vp]
    L.sfe_bow_db_add.argtypes = [vp, vp, i32, vp]
    L.sfe_bow_db_query.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    L.sfe_bow_db_size.argtypes = [vp]
    L.sfe_bow_transform_device.argtypes = [vp, vp, i32, vp, vp]
    L.sfe_kfdb_create.argtypes = [vp, vp, i32, C.POINTER(vp)]
    L.sfe_kfdb_destroy.argtypes = [vp]; L.sfe_kfdb_destroy.restype = None
    L.sfe_kfdb_size.argtypes = [vp]
    L.sfe_kfdb_reset.argtypes = [vp]
    L.sfe_kfdb_process_batch.argtypes = [vp, i32]
    L.sfe_kfdb_process_batch_on.argtypes = [vp, vp, i32]
    L.sfe_kfdb_results.argtypes = [vp] + [C.POINTER(vp)] * 6
    L.sfe_kfdb_submit_results.argtypes = [vp, C.POINTER(i32)]
    L.sfe_kfdb_wait_results.argtypes = [vp, i32] + [C.POINTER(vp)] * 6 + [C.POINTER(i32)]
    L.sfe_kfdb_store_pointers.argtypes = [vp, vp]
    L.sfe_kfdb_set_peers.argtypes = [vp, i32, i32, vp, vp]
    L.sfe_kfdb_insert_batch_on.argtypes = [vp, vp, i32]
    L.sfe_kfdb_query_batch_on.argtypes = [vp, vp, i32, i32]
    L.sfe_last_d2h_bytes.argtypes = [vp]; L.sfe_last_d2h_bytes.restype = C.c_int64
    L.sfe_d2h_refetches.argtypes = [vp]; L.sfe_d2h_refetches.restype = C.c_int64
    L.sfe_ipc_export.argtypes = [vp, vp, vp]; L.sfe_ipc_import.argtypes = [vp, vp, C.POINTER(vp)]; L.sfe_ipc_close.argtypes = [vp, vp]
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def default_config(**kw):
    cfg = SfeConfig()
    load_library().sfe_default_config(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise TypeError("unknown config field " + k)
        setattr(cfg, k, v)
    return cfg


class StereoFrontEnd:
    """One context bound to one GPU.  See module docstring for the reference members it mirrors."""

    def __init__(self, **cfg):
        self.lib = load_library()
        self.cfg = default_config(**cfg)
        h = C.c_void_p()
        r = self.lib.sfe_create(C.byref(self.cfg), C.byref(h))
        if r != SFE_OK:
            raise SfeError(r, self.lib.sfe_last_error(None).decode())
        self.h = h
        self.kp_cap = self.lib.sfe_keypoint_capacity(self.h)
        self.nlevels = self.lib.sfe_num_levels(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.sfe_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, r):
        if r != SFE_OK:
            raise SfeError(r, self.lib.sfe_last_error(self.h).decode())

    # --- thresholds (EST.cpp:275-311) ---
    def set_fast_threshold(self, th): self._chk(self.lib.sfe_set_fast_threshold(self.h, th))
    def get_fast_threshold(self): return self.lib.sfe_get_fast_threshold(self.h)
    def set_orb_threshold(self, th): self._chk(self.lib.sfe_set_orb_threshold(self.h, th))
    def get_orb_threshold(self): return self.lib.sfe_get_orb_threshold(self.h)

    def level_size(self, l):
        w, h, p, q = (C.c_int32() for _ in range(4))
        self._chk(self.lib.sfe_level_size(self.h, l, C.byref(w), C.byref(h), C.byref(p), C.byref(q)))
        return w.value, h.value, p.value, q.value

    @property
    def stream(self):
        return self.lib.sfe_stream(self.h)

    # --- per-frame stage ---
    def process_stereo(self, left, right, repeat=False):
        if repeat:
            self._chk(self.lib.sfe_process_stereo(self.h, None, None, 0, 1))
            return
        left = np.ascontiguousarray(left, np.uint8); right = np.ascontiguousarray(right, np.uint8)
        assert left.shape == right.shape == (self.cfg.height, self.cfg.width)
        self._chk(self.lib.sfe_process_stereo(self.h, _p(left), _p(right), left.shape[1], 0))

    def process_batch_host(self, images):
        """images: uint8 [npairs, 2, H, W] (or [2*npairs, H, W])."""
        images = np.ascontiguousarray(images, np.uint8).reshape(-1, self.cfg.height, self.cfg.width)
        npairs = images.shape[0] // 2
        self._chk(self.lib.sfe_process_batch_host(self.h, _p(images), npairs, self.cfg.width, self.cfg.width * self.cfg.height))
        return npairs

    def host_staging(self):
        n = C.c_size_t()
        p = self.lib.sfe_host_staging(self.h, C.byref(n))
        buf = (C.c_uint8 * n.value).from_address(p)
        return np.frombuffer(buf, np.uint8).reshape(-1, self.cfg.height, self.cfg.width)

    def process_staged(self, npairs):
        """H2D from the context's pinned staging + kernels, synchronous (the e2e path)."""
        n = C.c_size_t()
        p = self.lib.sfe_host_staging(self.h, C.byref(n))
        self._chk(self.lib.sfe_process_batch_host(self.h, p, npairs, self.cfg.width, self.cfg.width * self.cfg.height))

    def submit_staged(self, npairs):
        """Asynchronous: H2D from pinned staging + kernels + D2H of the results (sfe_submit_batch_host)."""
        n = C.c_size_t()
        p = self.lib.sfe_host_staging(self.h, C.byref(n))
        self._chk(self.lib.sfe_submit_batch_host(self.h, p, npairs, self.cfg.width, self.cfg.width * self.cfg.height))

    def wait_batch(self):
        """Block until the submitted batch's results are in pinned host memory; returns numpy views
        (keypoints [2*npairs, cap], descriptors [2*npairs, cap, 32], matches [npairs, cap], counts [npairs])."""
        ptrs = [C.c_void_p() for _ in range(4)]
        self._chk(self.lib.sfe_wait_batch(self.h, *[C.byref(p) for p in ptrs]))
        return ptrs

    def upload(self, images):
        images = np.ascontiguousarray(images, np.uint8).reshape(-1, self.cfg.height, self.cfg.width)
        npairs = images.shape[0] // 2
        self._chk(self.lib.sfe_upload_batch(self.h, _p(images), npairs, self.cfg.width, self.cfg.width * self.cfg.height))
        self._chk(self.lib.sfe_synchronize(self.h))
        return npairs

    def upload_async(self, npairs):
        """H2D of the context's own pinned staging buffer, no synchronisation (link probes)."""
        n = C.c_size_t()
        p = self.lib.sfe_host_staging(self.h, C.byref(n))
        self._chk(self.lib.sfe_upload_batch(self.h, p, npairs, self.cfg.width, self.cfg.width * self.cfg.height))

    def process_device(self, npairs):
        self._chk(self.lib.sfe_process_batch_device(self.h, npairs))

    def synchronize(self):
        self._chk(self.lib.sfe_synchronize(self.h))

    # --- results (CStereoSLAMKF.h:68-70) ---
    def get_values(self, pair=0):
        cap = self.kp_cap
        kl = np.zeros(cap, KEYPOINT_DTYPE); kr = np.zeros(cap, KEYPOINT_DTYPE)
        dl = np.zeros((cap, 32), np.uint8); dr = np.zeros((cap, 32), np.uint8)
        m = np.zeros(cap, DMATCH_DTYPE); cnt = np.zeros(1, COUNTS_DTYPE)
        self._chk(self.lib.sfe_get_values(self.h, pair, _p(kl), _p(kr), _p(dl), _p(dr), _p(m), _p(cnt)))
        nl, nr, nm = int(cnt[0]["n_left"]), int(cnt[0]["n_right"]), int(cnt[0]["n_matches"])
        return kl[:nl], kr[:nr], dl[:nl], dr[:nr], m[:nm]

    def download_batch(self):
        ptrs = [C.c_void_p() for _ in range(4)]
        self._chk(self.lib.sfe_download_batch(self.h, *[C.byref(p) for p in ptrs]))
        return ptrs

    def get_level(self, img, level, blurred=False):
        w, h, _, _ = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        self._chk(self.lib.sfe_get_level(self.h, img, level, int(blurred), _p(out)))
        return out

    def get_raw_corners(self, img, level):
        w, h, _, _ = self.level_size(level)
        cap = (w // 2 + 1) * (h // 2 + 1)
        out = np.zeros(cap, np.uint32); n = C.c_int32()
        self._chk(self.lib.sfe_get_raw_corners(self.h, img, level, _p(out), cap, C.byref(n)))
        r = out[:n.value]
        return (r & 0xfff).astype(np.int32), ((r >> 12) & 0xfff).astype(np.int32), (r >> 24).astype(np.int32)

    # --- brute-force match (EST.cpp:1447-1452) ---
    def match_bf(self, query, train):
        q = np.ascontiguousarray(query, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
        out = np.zeros(max(len(q), 1), DMATCH_DTYPE); n = C.c_int32()
        self._chk(self.lib.sfe_match_bf(self.h, _p(q), len(q), _p(t), len(t), _p(out), C.byref(n)))
        return out[:n.value]

    # --- timing ---
    def timer_start(self): self._chk(self.lib.sfe_timer_start(self.h))
    def stream_wait(self, other): self._chk(self.lib.sfe_stream_wait(self.h, other.h))
    def timer_stop(self):
        ms = C.c_float()
        self._chk(self.lib.sfe_timer_stop(self.h, C.byref(ms)))
        return ms.value
    def set_profiling(self, on): self._chk(self.lib.sfe_set_profiling(self.h, int(on)))
    def stage_times(self):
        n = self.lib.sfe_num_stages()
        ms = np.zeros(n, np.float32); la = np.zeros(n, np.int32)
        self._chk(self.lib.sfe_stage_times(self.h, _p(ms), _p(la)))
        return {self.lib.sfe_stage_name(i).decode(): (float(ms[i]), int(la[i])) for i in range(n)}
    def kernel_launches(self): return int(self.lib.sfe_kernel_launches(self.h))
    def last_d2h_bytes(self): return int(self.lib.sfe_last_d2h_bytes(self.h))
    def d2h_refetches(self): return int(self.lib.sfe_d2h_refetches(self.h))


class BowDatabase:
    """DBoW2 BriefVocabulary + BriefDatabase on the GPU (CBoWManager.h:44-111).  `voc` is the flattened
    vocabulary dict produced by voc_loader.load_vocabulary()."""

    def __init__(self, fe, voc):
        self.fe = fe; self.lib = fe.lib
        h = C.c_void_p()
        arrs = [np.ascontiguousarray(voc["child_ofs"], np.int32), np.ascontiguousarray(voc["child_idx"], np.int32),
                np.ascontiguousarray(voc["node_desc"], np.uint8), np.ascontiguousarray(voc["node_weight"], np.float64),
                np.ascontiguousarray(voc["node_word"], np.int32)]
        fe._chk(self.lib.sfe_bow_create(fe.h, int(voc["nnodes"]), int(voc["nwords"]), *[_p(a) for a in arrs], C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.sfe_bow_destroy(self.h); self.h = None

    def transform(self, desc):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        w = np.zeros(len(d), np.int32); wt = np.zeros(len(d), np.float64)
        self.fe._chk(self.lib.sfe_bow_transform(self.h, _p(d), len(d), _p(w), _p(wt)))
        return w, wt

    def vector(self, desc):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        w = np.zeros(max(len(d), 1), np.int32); v = np.zeros(max(len(d), 1), np.float64); n = C.c_int32()
        self.fe._chk(self.lib.sfe_bow_vector(self.h, _p(d), len(d), _p(w), _p(v), C.byref(n)))
        return w[:n.value], v[:n.value]

    def add(self, desc):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); eid = C.c_int32()
        self.fe._chk(self.lib.sfe_bow_db_add(self.h, _p(d), len(d), C.byref(eid)))
        return eid.value

    def query(self, desc, max_results=4):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        ids = np.zeros(max_results, np.int32); sc = np.zeros(max_results, np.float64); n = C.c_int32()
        self.fe._chk(self.lib.sfe_bow_db_query(self.h, _p(d), len(d), max_results, _p(ids), _p(sc), C.byref(n)))
        return [(int(ids[i]), float(sc[i])) for i in range(n.value)]

    def size(self): return self.lib.sfe_bow_db_size(self.h)


class KeyframeDatabase:
    """Batched full front-end (sfe_kfdb_*): BoW query/add + keyframe store + data association for every pair of the
    last processed stereo batch -- the per-keyframe sequence of EST.cpp:413-447 and :919, device-resident."""

    def __init__(self, fe, bow, capacity_frames):
        self.fe = fe; self.lib = fe.lib; self.bow = bow
        h = C.c_void_p()
        fe._chk(self.lib.sfe_kfdb_create(fe.h, bow.h, int(capacity_frames), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.sfe_kfdb_destroy(self.h); self.h = None

    def size(self): return self.lib.sfe_kfdb_size(self.h)
    def reset(self): self.fe._chk(self.lib.sfe_kfdb_reset(self.h))
    def process_batch(self, npairs, src=None):
        """src: the StereoFrontEnd whose last batch is consumed (default: the one the database was created on)."""
        if src is None:
            self.fe._chk(self.lib.sfe_kfdb_process_batch(self.h, npairs))
        else:
            src._chk(self.lib.sfe_kfdb_process_batch_on(self.h, src.h, npairs))

    def _views(self, p, n):
        cap = self.fe.kp_cap

        def view(ptr, dtype, shape):
            cnt = int(np.prod(shape))
            buf = (C.c_uint8 * (cnt * np.dtype(dtype).itemsize)).from_address(ptr.value)
            return np.frombuffer(buf, dtype, cnt).reshape(shape).copy()
        return {"query_ids": view(p[0], np.int32, (n, 4)), "query_scores": view(p[1], np.float64, (n, 4)),
                "query_n": view(p[2], np.int32, (n,)), "pair_kf": view(p[3], np.int32, (n, 5)),
                "best": view(p[4], np.uint32, (n, 5, cap)), "matched_n": view(p[5], np.int32, (n,))}

    def results(self, npairs):
        """Returns dict of numpy arrays: query_ids [n,4], query_scores [n,4], query_n [n], pair_kf [n,5],
        best [n,5,cap] (packed dist<<16 | train index, 0xFFFFFFFF = none), matched_n [n]."""
        p = [C.c_void_p() for _ in range(6)]
        self.fe._chk(self.lib.sfe_kfdb_results(self.h, *[C.byref(x) for x in p]))
        return self._views(p, npairs)

    # --- streaming results (pinned slots; `best` rows are only valid up to matched_n) ---
    def submit_results(self):
        s = C.c_int32()
        self.fe._chk(self.lib.sfe_kfdb_submit_results(self.h, C.byref(s)))
        return s.value

    def wait_results(self, slot, copy=True):
        p = [C.c_void_p() for _ in range(6)]; n = C.c_int32()
        self.fe._chk(self.lib.sfe_kfdb_wait_results(self.h, slot, *[C.byref(x) for x in p], C.byref(n)))
        if not copy:                       # the caller only needs the small arrays (bench): view query_n without copying `best`
            buf = (C.c_int32 * n.value).from_address(p[2].value)
            return {"query_n": np.frombuffer(buf, np.int32, n.value), "n": n.value}
        return self._views(p, n.value)

    # --- two-pass form + one logical database over several ranks ---
    def insert_batch(self, npairs, src=None):
        src = src or self.fe
        src._chk(self.lib.sfe_kfdb_insert_batch_on(self.h, src.h, npairs))

    def query_batch(self, first_local_frame, npairs, src=None):
        src = src or self.fe
        src._chk(self.lib.sfe_kfdb_query_batch_on(self.h, src.h, first_local_frame, npairs))

    def store_pointers(self):
        out = (C.c_void_p * 5)()
        self.fe._chk(self.lib.sfe_kfdb_store_pointers(self.h, out))
        return [int(x or 0) for x in out]

    def set_peers(self, rank, starts, ptr_rows):
        """starts: nranks+1 global shard starts; ptr_rows: nranks rows of 5 device pointers (own row ignored)."""
        nr = len(ptr_rows)
        st = np.ascontiguousarray(starts, np.int32); pt = np.ascontiguousarray(np.array(ptr_rows, np.uint64).reshape(nr, 5))
        self._peer_keep = (st, pt)
        self.fe._chk(self.lib.sfe_kfdb_set_peers(self.h, nr, rank, _p(st), _p(pt)))

    def connect(self, dist, frames_per_rank):
        """One logical database over the ranks of `dist` (torch.distributed, one process per GPU): exports this rank's five store
        arrays with CUDA IPC, maps the peers', tells the kernels.  frames_per_rank: list of shard sizes (global order = rank order)."""
        rank, world = dist.get_rank(), dist.get_world_size()
        mine = self.store_pointers()
        handles = []
        for ptr in mine:
            h = (C.c_uint8 * 64)()
            self.fe._chk(self.lib.sfe_ipc_export(self.fe.h, ptr, h)); handles.append(bytes(h))
        gathered = [None] * world
        dist.all_gather_object(gathered, handles)
        rows = []; self._imports = []
        for r in range(world):
            if r == rank:
                rows.append(mine); continue
            row = []
            for hb in gathered[r]:
                p = C.c_void_p(); hbuf = (C.c_uint8 * 64).from_buffer_copy(hb)
                self.fe._chk(self.lib.sfe_ipc_import(self.fe.h, hbuf, C.byref(p)))
                self._imports.append(p.value); row.append(p.value)
            rows.append(row)
        starts = np.concatenate([[0], np.cumsum(frames_per_rank)]).astype(np.int32)
        self.set_peers(rank, starts, rows)
        return starts
