#!/usr/bin/env python
"""Loop-closure query batch (BASELINE.json configs[3]): NQ query keyframes vs an NDB-keyframe descriptor database
sharded over the GPUs of one box; 5 candidates per query (the planted revisit + 4 random database keyframes).
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/lc_bench.py
Compares the product path (peer loads through NVLink inside the match kernel) with the NCCL all-gather baseline."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ndb", type=int, default=20000); ap.add_argument("--nq", type=int, default=4096)
    ap.add_argument("--nmax", type=int, default=500); ap.add_argument("--k", type=int, default=5)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    d = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local)); d = dist
    from srba_stereo_slam_b200 import frontend, loop_closure as lc, sharding
    fe = frontend.StereoFrontEnd(width=320, height=240, nlevels=1, nfeatures=100, max_pairs=1, device=local)
    m = lc.LoopClosureMatcher(fe, a.nmax, d)
    s, e = sharding.shard_range(a.ndb, rank, world)
    dbc, dbb = lc.synth_keyframe_blocks(e - s, a.nmax, seed=100 + rank)
    qs, qe = sharding.shard_range(a.nq, rank, world)
    rng = np.random.default_rng(7 + rank)
    src_local = rng.integers(0, e - s, qe - qs)                       # planted revisit: a keyframe of the local shard
    qc, qb = lc.synth_keyframe_blocks(qe - qs, a.nmax, seed=200 + rank, planted_from=(dbb, dbc, src_local))
    cand = rng.integers(0, a.ndb, (qe - qs, a.k)).astype(np.int32)
    cand[:, 0] = s + src_local
    m.set_database(dbc, dbb, a.ndb)
    m.upload_queries(qc, qb)

    def timed(mode, with_gather):
        ts = []
        for _ in range(a.reps + 1):
            if d is not None:
                d.barrier()
            torch.cuda.synchronize(); fe.synchronize()
            t0 = time.perf_counter()
            if with_gather:
                m.replicate_database_allgather()
            m.run(mode); fe.synchronize()
            if d is not None:
                d.barrier()
            ts.append(time.perf_counter() - t0)
        t = torch.tensor([min(ts[1:])], device="cuda")
        if d is not None:
            d.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    m.build_pairs(cand, "peer")
    t_peer = timed("peer", False)
    best_peer = m.results()
    m.replicate_database_allgather()
    m.build_pairs(cand, "allgather")
    t_ag_match = timed("allgather", False)
    t_ag_total = timed("allgather", True)
    best_ag = m.results()
    same = bool(np.array_equal(best_peer, best_ag))
    hits = int(((best_peer[:, 0, :] >> 16) <= 30).sum())
    remote = float((((cand < s) | (cand >= e)) & (cand >= 0)).mean())
    if rank == 0:
        pairs = a.nq * a.k
        popc32 = float(pairs) * a.nmax * a.nmax * 8
        print(json.dumps({"workload": "configs[3]: %d query KFs x %d candidates vs %d-KF database, %d desc/KF" % (a.nq, a.k, a.ndb, a.nmax),
                          "n_gpus": world, "peer_ms": t_peer * 1e3, "allgather_match_ms": t_ag_match * 1e3, "allgather_total_ms": t_ag_total * 1e3,
                          "query_kf_per_s_peer": a.nq / t_peer, "query_kf_per_s_allgather": a.nq / t_ag_total,
                          "equivalent_popc32_per_s_peer": popc32 / t_peer, "results_identical": same, "planted_hits_rank0": hits,
                          "remote_candidate_fraction_rank0": remote,
                          "db_bytes_total": a.ndb * a.nmax * 32, "timing": "host wall clock, barrier + sync both sides, min of %d, max over ranks" % a.reps}))
    m.close(); fe.close()
    if d is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
