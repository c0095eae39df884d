"""Loop-closure query batches over a keyframe database sharded across the GPUs of one box
(BASELINE.json configs[3]: 4096 query keyframes vs a 20 000-keyframe descriptor database on 8 B200).

Each rank owns a contiguous shard of the database (fixed-stride blocks {n; desc[nmax][32]}) and a
contiguous shard of the query keyframes.  For every query the candidate database keyframes (previous keyframe +
BoW hits, EST.cpp:435, :1758) are matched brute force (BFMatcher semantics, EST.cpp:1447-1452).

Two exchange strategies for the database side:
  * `peer`      -- the product path: shards are exported with CUDA IPC, every rank maps its peers' shards and the
                   match kernel loads candidate descriptors straight through NVLink (sfe_match_bf_pairs_device);
                   only the blocks that are matched cross the link and the transfer overlaps the popc work.
  * `allgather` -- the baseline: one NCCL all-gather replicates the whole database, then the same kernel runs on
                   local memory.
PyTorch is used for the process group / NCCL only; all matching is the CUDA library's.
"""
import ctypes as C
import numpy as np

from . import frontend, sharding

BF_PAIR_DTYPE = np.dtype([("q_start", "<i4"), ("nq", "<i4"), ("t_rank", "<i4"), ("t_start", "<i4"), ("nt", "<i4"), ("out_start", "<i4")])


class LoopClosureMatcher:
    def __init__(self, fe, nmax, dist=None):
        self.fe, self.lib, self.nmax, self.dist = fe, fe.lib, int(nmax), dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        for n in ("sfe_device_alloc", "sfe_device_free", "sfe_copy_h2d", "sfe_copy_d2h", "sfe_ipc_export", "sfe_ipc_import", "sfe_ipc_close",
                  "sfe_match_bf_pairs_device"):
            getattr(self.lib, n).restype = C.c_int
        self.lib.sfe_device_alloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        self.lib.sfe_device_free.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.sfe_copy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        self.lib.sfe_copy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        self.lib.sfe_ipc_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.sfe_ipc_import.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        self.lib.sfe_ipc_close.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.sfe_match_bf_pairs_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t]
        self._allocs, self._imports = [], []
        self.n_total = 0; self.db_counts = None; self.d_bases = None; self.d_bases_local = None

    # ---- raw device memory helpers --------------------------------------------------------------------------------
    def _alloc(self, nbytes):
        p = C.c_void_p()
        self.fe._chk(self.lib.sfe_device_alloc(self.fe.h, max(int(nbytes), 256), C.byref(p)))
        self._allocs.append(p.value)
        return p.value

    def _h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self.fe._chk(self.lib.sfe_copy_h2d(self.fe.h, dptr, arr.ctypes.data_as(C.c_void_p), arr.nbytes))

    def close(self):
        for p in self._imports:
            self.lib.sfe_ipc_close(self.fe.h, p)
        for p in self._allocs:
            self.lib.sfe_device_free(self.fe.h, p)
        self._allocs, self._imports = [], []

    # ---- database ---------------------------------------------------------------------------------------------------
    def set_database(self, counts_local, blocks_local, n_total):
        """counts_local [k] int32, blocks_local [k, nmax, 32] uint8: this rank's shard (shard_range(n_total, rank, world))."""
        s, e = sharding.shard_range(n_total, self.rank, self.world)
        assert len(counts_local) == e - s and blocks_local.shape == (e - s, self.nmax, 32)
        self.n_total = n_total
        self.d_db_local = self._alloc(blocks_local.nbytes)
        self._h2d(self.d_db_local, blocks_local)
        self.blocks_local_host = blocks_local
        if self.dist is None:
            self.db_counts = np.asarray(counts_local, np.int32)
            bases = np.array([self.d_db_local], np.uint64)
        else:
            import torch
            handle = (C.c_uint8 * 64)()
            self.fe._chk(self.lib.sfe_ipc_export(self.fe.h, self.d_db_local, handle))
            gathered = [None] * self.world
            self.dist.all_gather_object(gathered, (bytes(handle), np.asarray(counts_local, np.int32)))
            bases = np.zeros(self.world, np.uint64)
            for r, (h, _) in enumerate(gathered):
                if r == self.rank:
                    bases[r] = self.d_db_local
                else:
                    p = C.c_void_p()
                    hb = (C.c_uint8 * 64).from_buffer_copy(h)
                    self.fe._chk(self.lib.sfe_ipc_import(self.fe.h, hb, C.byref(p)))
                    self._imports.append(p.value); bases[r] = p.value
            self.db_counts = np.concatenate([c for _, c in gathered]).astype(np.int32)
        self.d_bases = self._alloc(bases.nbytes); self._h2d(self.d_bases, bases)

    def replicate_database_allgather(self):
        """Baseline: NCCL all-gather of the whole database; returns the device tensor (keeps it alive) and sets the
        single-base pointer table used by match(..., mode='allgather')."""
        import torch
        dev = torch.device("cuda", self.fe.cfg.device)
        if getattr(self, "_local_t", None) is None:
            self._local_t = torch.from_numpy(self.blocks_local_host).to(dev)      # resident shard: only the exchange is timed
        local = self._local_t
        if self.dist is None:
            full = local
        else:
            kmax = max(e - s for s, e in (sharding.shard_range(self.n_total, r, self.world) for r in range(self.world)))
            pad = torch.zeros((kmax, self.nmax, 32), dtype=torch.uint8, device=dev); pad[:local.shape[0]] = local
            allb = torch.empty((self.world, kmax, self.nmax, 32), dtype=torch.uint8, device=dev)
            self.dist.all_gather_into_tensor(allb.view(-1), pad.view(-1))
            full = torch.cat([allb[r, :e - s] for r, (s, e) in enumerate(sharding.shard_range(self.n_total, r, self.world) for r in range(self.world))])
        torch.cuda.synchronize(dev)
        self.full_db = full
        b = np.array([full.data_ptr()], np.uint64)
        self.d_bases_local = self._alloc(8); self._h2d(self.d_bases_local, b)
        return full

    # ---- matching -----------------------------------------------------------------------------------------------------
    def upload_queries(self, q_counts, q_blocks):
        self.q_counts = np.asarray(q_counts, np.int32)
        self.d_q = self._alloc(q_blocks.nbytes); self._h2d(self.d_q, q_blocks)

    def build_pairs(self, candidates, mode="peer"):
        """candidates: int32 [nq, K] global database keyframe ids (-1 = empty slot)."""
        cand = np.asarray(candidates, np.int32)
        nq, K = cand.shape
        starts = np.array([sharding.shard_range(self.n_total, r, self.world)[0] for r in range(self.world)] + [self.n_total])
        pairs = np.zeros(nq * K, BF_PAIR_DTYPE)
        valid = cand.ravel() >= 0
        ids = np.where(valid, cand.ravel(), 0)
        owner = np.searchsorted(starts, ids, side="right") - 1
        pairs["q_start"] = np.repeat(np.arange(nq, dtype=np.int32) * self.nmax, K)
        pairs["nq"] = np.where(valid, np.repeat(self.q_counts, K), 0)
        pairs["nt"] = np.where(valid, self.db_counts[ids], 0)
        if mode == "peer":
            pairs["t_rank"] = owner; pairs["t_start"] = (ids - starts[owner]) * self.nmax
        else:
            pairs["t_rank"] = 0; pairs["t_start"] = ids * self.nmax
        pairs["out_start"] = np.arange(nq * K, dtype=np.int32) * self.nmax
        self.npairs, self.K, self.nq = nq * K, K, nq
        self.d_pairs = self._alloc(pairs.nbytes); self._h2d(self.d_pairs, pairs)
        self.d_best = self._alloc(nq * K * self.nmax * 4)
        return pairs

    def run(self, mode="peer"):
        """Asynchronous on the front-end's stream."""
        bases = self.d_bases if mode == "peer" else self.d_bases_local
        self.fe._chk(self.lib.sfe_match_bf_pairs_device(self.fe.h, self.d_q, bases, self.d_pairs, self.npairs, self.nmax,
                                                        self.d_best, self.npairs * self.nmax))

    def results(self):
        out = np.zeros((self.nq, self.K, self.nmax), np.uint32)
        self.fe._chk(self.lib.sfe_copy_d2h(self.fe.h, out.ctypes.data_as(C.c_void_p), self.d_best, out.nbytes))
        return out


def synth_keyframe_blocks(n_kf, nmax, seed, n_min=None, planted_from=None, planted_frac=0.5, max_flip=30):
    """Random 256-bit descriptors; with `planted_from` (blocks, counts) a fraction of each keyframe's descriptors are
    near-duplicates (<= max_flip flipped bits) of the same-index keyframe of that set, so matches are non-trivial."""
    rng = np.random.default_rng(seed)
    counts = rng.integers(n_min if n_min else nmax, nmax + 1, n_kf).astype(np.int32)
    blocks = rng.integers(0, 256, (n_kf, nmax, 32), dtype=np.uint8)
    for i in range(n_kf):
        blocks[i, counts[i]:] = 0
    if planted_from is not None:
        src_b, src_c, src_ids = planted_from
        for i in range(n_kf):
            j = src_ids[i]
            m = int(min(counts[i], src_c[j]) * planted_frac)
            sel = rng.permutation(int(src_c[j]))[:m]
            d = src_b[j, sel].copy()
            flips = rng.integers(0, 256, (m, max_flip))
            rows = np.repeat(np.arange(m), max_flip)
            np.bitwise_xor.at(d, (rows, (flips >> 3).ravel()), (1 << (flips & 7)).astype(np.uint8).ravel())
            blocks[i, :m] = d
    return counts, blocks
