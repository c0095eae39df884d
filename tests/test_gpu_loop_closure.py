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


def test_lc_pairs_at_benchmark_size():
    """configs[3] shape: 500 descriptors per keyframe (ragged 350..500), 96 queries x 5 candidates vs a 400-keyframe database --
    every one of the 480 keyframe pairs against the oracle's brute force (lowest-index ties, float distance)."""
    from srba_stereo_slam_b200 import frontend, loop_closure as lc
    from oracle import oracle as orc
    nmax, ndb, nq, K = 500, 400, 96, 5
    rng = np.random.default_rng(11)
    dbc, dbb = lc.synth_keyframe_blocks(ndb, nmax, seed=5, n_min=350)
    src = rng.integers(0, ndb, nq)
    qc, qb = lc.synth_keyframe_blocks(nq, nmax, seed=6, n_min=350, planted_from=(dbb, dbc, src))
    cand = rng.integers(0, ndb, (nq, K)).astype(np.int32); cand[:, 0] = src
    cand[rng.integers(0, nq, 10), rng.integers(1, K, 10)] = -1
    dbb[7, :dbc[7]] = dbb[7, 0]                                  # a keyframe of identical descriptors: every query ties on index 0
    cand[5, 1] = 7
    fe = frontend.StereoFrontEnd(width=320, height=240, nlevels=1, nfeatures=100, max_pairs=1)
    m = lc.LoopClosureMatcher(fe, nmax)
    try:
        m.set_database(dbc, dbb, ndb); m.upload_queries(qc, qb)
        m.build_pairs(cand, mode="peer"); m.run("peer")
        best = m.results()
        for i in range(nq):
            for k in range(K):
                if cand[i, k] < 0:
                    assert np.all(best[i, k] == 0xFFFFFFFF); continue
                o = orc.match_bf(qb[i, :qc[i]], dbb[cand[i, k], :dbc[cand[i, k]]])
                exp = (o["distance"].astype(np.uint32) << 16) | o["trainIdx"].astype(np.uint32)
                assert np.array_equal(best[i, k, :qc[i]], exp), (i, k)
                assert np.all(best[i, k, qc[i]:] == 0xFFFFFFFF)
        assert np.all((best[5, 1, :qc[5]] & 0xFFFF) == 0)
    finally:
        m.close(); fe.close()
