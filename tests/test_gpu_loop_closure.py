"""GPU (-m gpu): pair-list brute force over a (single-rank) keyframe database -- the matcher of the loop-closure batch
(BASELINE.json configs[3]) -- against the oracle; peer and all-gather modes must agree."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_lc_pairs_match_oracle():
    from srba_stereo_slam_b200 import frontend, loop_closure as lc
    from oracle import oracle as orc
    nmax, ndb, nq, K = 96, 40, 12, 5
    rng = np.random.default_rng(3)
    dbc, dbb = lc.synth_keyframe_blocks(ndb, nmax, seed=1, n_min=50)
    src = rng.integers(0, ndb, nq)
    qc, qb = lc.synth_keyframe_blocks(nq, nmax, seed=2, n_min=40, planted_from=(dbb, dbc, src))
    cand = rng.integers(0, ndb, (nq, K)).astype(np.int32)
    cand[:, 0] = src                       # the true revisit is always a candidate
    cand[3, 2] = -1; cand[7, 4] = -1       # empty slots
    fe = frontend.StereoFrontEnd(width=320, height=240, nlevels=1, nfeatures=100, max_pairs=1)
    m = lc.LoopClosureMatcher(fe, nmax)
    try:
        m.set_database(dbc, dbb, ndb)
        m.upload_queries(qc, qb)
        m.build_pairs(cand, mode="peer"); m.run("peer")
        best = m.results()
        m.replicate_database_allgather()
        m.build_pairs(cand, mode="allgather"); m.run("allgather")
        best2 = m.results()
        assert np.array_equal(best, best2)
        planted_hits = 0
        for i in range(nq):
            for k in range(K):
                if cand[i, k] < 0:
                    assert np.all(best[i, k] == 0xFFFFFFFF); continue
                o = orc.match_bf(qb[i, :qc[i]], dbb[cand[i, k], :dbc[cand[i, k]]])
                assert np.array_equal(best[i, k, :qc[i]] & 0xFFFF, o["trainIdx"].astype(np.uint32))
                assert np.array_equal(best[i, k, :qc[i]] >> 16, o["distance"].astype(np.uint32))
                assert np.all(best[i, k, qc[i]:] == 0xFFFFFFFF)
            planted_hits += int(((best[i, 0, :qc[i]] >> 16) <= 30).sum())
        assert planted_hits > nq * 10
    finally:
        m.close(); fe.close()
