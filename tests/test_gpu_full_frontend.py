"""GPU (-m gpu): the batched full front-end (BASELINE.json configs[2] shape): extract + L/R match, then per frame BoW
query (top-4 among OLDER keyframes) + add, matched-subset gather, brute-force association against the previous keyframe
and the BoW hits -- against the oracle run frame by frame in the reference's order (EST.cpp:413-447, :919)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_full_frontend_batches_match_sequential_oracle(voc):
    from srba_stereo_slam_b200 import frontend, synth
    from oracle import oracle as orc
    W, H, NL, NFEAT, NF, B = 640, 192, 4, 300, 7, 3
    frames = synth.CorridorScene(seed=8, width=W, height=H).sequence(NF)
    fe = frontend.StereoFrontEnd(width=W, height=H, nlevels=NL, nfeatures=NFEAT, max_pairs=B)
    bow = frontend.BowDatabase(fe, voc)
    db = frontend.KeyframeDatabase(fe, bow, capacity_frames=16)
    try:
        got = []
        for s in range(0, NF, B):
            n = min(B, NF - s)
            fe.process_batch_host(frames[s:s + n])
            db.process_batch(n)
            r = db.results(n)
            for i in range(n):
                got.append({k: v[i] for k, v in r.items()})
        assert db.size() == NF
        # ---- oracle, one frame at a time
        odb, msubs = [], []
        for g in range(NF):
            ol, odl, _ = orc.orb_extract(frames[g, 0], nfeatures=NFEAT, nlevels=NL)
            orr, odr, _ = orc.orb_extract(frames[g, 1], nfeatures=NFEAT, nlevels=NL)
            om = orc.match_rowband(ol, odl, orr, odr)
            cur = odl[om["queryIdx"]]
            res = orc.bow_query(voc, odb, odl, 4)                       # query BEFORE add (EST.cpp:424 vs :919)
            out = got[g]
            assert out["matched_n"] == len(cur)
            assert out["query_n"] == len(res)
            assert out["query_ids"][:len(res)].tolist() == [r_[0] for r_ in res]
            assert out["query_scores"][:len(res)].tolist() == [r_[1] for r_ in res]      # FP64, bit-identical
            exp_kf = [g - 1 if g > 0 else -1] + [(r_[0] if r_[0] != g - 1 else -1) for r_ in res] + [-1] * (4 - len(res))
            assert out["pair_kf"].tolist() == exp_kf
            for slot, kf in enumerate(exp_kf):
                best = out["best"][slot][:len(cur)]
                if kf < 0:
                    assert np.all(out["best"][slot] == 0xFFFFFFFF)
                    continue
                obf = orc.match_bf(cur, msubs[kf])
                assert np.array_equal(best & 0xFFFF, obf["trainIdx"].astype(np.uint32)), (g, slot)
                assert np.array_equal(best >> 16, obf["distance"].astype(np.uint32)), (g, slot)
                assert np.all(out["best"][slot][len(cur):] == 0xFFFFFFFF)
            w_, wt_ = orc.bow_transform(voc, odl)
            odb.append(orc.bow_vector(voc, w_, wt_)); msubs.append(cur)
        assert any(o["query_n"] == 4 for o in got[4:])
    finally:
        db.close(); bow.close(); fe.close()


def _oracle_frontend(voc, frames, nfeat, nl):
    """The sequential oracle of the batched full front-end: per frame (keypoint counts, matched subset, query top-4, pair list, BF keys)."""
    from oracle import oracle as orc
    odb, msubs, out = [], [], []
    for g in range(len(frames)):
        ol, odl, _ = orc.orb_extract(frames[g, 0], nfeatures=nfeat, nlevels=nl)
        orr, odr, _ = orc.orb_extract(frames[g, 1], nfeatures=nfeat, nlevels=nl)
        om = orc.match_rowband(ol, odl, orr, odr)
        cur = odl[om["queryIdx"]]
        res = orc.bow_query(voc, odb, odl, 4)
        kf = [g - 1 if g > 0 else -1] + [(r_[0] if r_[0] != g - 1 else -1) for r_ in res] + [-1] * (4 - len(res))
        keys = []
        for k in kf:
            if k < 0:
                keys.append(None); continue
            o = orc.match_bf(cur, msubs[k])
            keys.append((o["distance"].astype(np.uint32) << 16) | o["trainIdx"].astype(np.uint32))
        out.append(dict(n=len(cur), res=res, kf=kf, keys=keys))
        w_, wt_ = orc.bow_transform(voc, odl)
        odb.append(orc.bow_vector(voc, w_, wt_)); msubs.append(cur)
    return out


def _check_frame(got, exp, tag):
    n = exp["n"]
    assert got["matched_n"] == n, tag
    assert got["query_n"] == len(exp["res"]), tag
    assert got["query_ids"][:len(exp["res"])].tolist() == [r_[0] for r_ in exp["res"]], tag
    assert got["query_scores"][:len(exp["res"])].tolist() == [r_[1] for r_ in exp["res"]], tag
    assert got["pair_kf"].tolist() == exp["kf"], tag
    for slot, k in enumerate(exp["kf"]):
        if k is None or exp["keys"][slot] is None or (len(exp["keys"][slot]) == 0 and n > 0):      # no keyframe in the slot / it has no stereo matches
            assert np.all(got["best"][slot][:max(n, 1)] == 0xFFFFFFFF), (tag, slot)
        else:
            assert np.array_equal(got["best"][slot][:n], exp["keys"][slot]), (tag, slot)


def test_one_logical_database_over_ranks_equals_single_rank(voc):
    """BASELINE.json configs[2]: frames sharded over G ranks, ONE keyframe database (the reference has one: CBoWManager.h:71-88,
    EST.cpp:424-427, :1432-1452).  G = 3 ranks emulated as three contexts + stores on one GPU wired together with
    sfe_kfdb_set_peers (on a real box the pointers come through CUDA IPC: KeyframeDatabase.connect, exercised by bench.py --gpus N):
    two passes (insert everything, then query) must reproduce, bit for bit, what one rank processing the whole sequence reports --
    cross-shard BoW hits, previous-keyframe links across the shard boundary and brute-force keys against peer stores included."""
    from srba_stereo_slam_b200 import frontend, synth
    W, H, NL, NFEAT, B = 640, 192, 4, 300, 4
    shards = [8, 5, 7]                                    # ragged on purpose
    NF = sum(shards)
    seq = synth.CorridorScene(seed=8, width=W, height=H).sequence(8)
    frames = np.concatenate([seq[0:8], seq[3:8], seq[0:7]])                  # ranks 1 and 2 REVISIT places rank 0 owns: cross-shard loop closures
    assert len(frames) == NF
    exp = _oracle_frontend(voc, frames, NFEAT, NL)
    # ---- single rank, streaming (insert + query per batch) ----
    fe = frontend.StereoFrontEnd(width=W, height=H, nlevels=NL, nfeatures=NFEAT, max_pairs=B)
    bow = frontend.BowDatabase(fe, voc); db = frontend.KeyframeDatabase(fe, bow, capacity_frames=32)
    single = []
    try:
        for s in range(0, NF, B):
            n = min(B, NF - s)
            fe.process_batch_host(frames[s:s + n]); db.process_batch(n)
            slot = db.submit_results(); r = db.wait_results(slot)              # the streaming download path
            r2 = db.results(n)
            for k in r:
                if k != "best":
                    assert np.array_equal(r[k], r2[k]), k
            for i in range(n):
                assert np.array_equal(r["best"][i][:, :r["matched_n"][i]], r2["best"][i][:, :r["matched_n"][i]])
                single.append({k: v[i] for k, v in r2.items()})
    finally:
        db.close(); bow.close(); fe.close()
    for g in range(NF):
        _check_frame(single[g], exp[g], ("single", g))
    # ---- three ranks, two passes ----
    fes = [frontend.StereoFrontEnd(width=W, height=H, nlevels=NL, nfeatures=NFEAT, max_pairs=B) for _ in shards]
    bows = [frontend.BowDatabase(f, voc) for f in fes]
    dbs = [frontend.KeyframeDatabase(f, b, capacity_frames=10) for f, b in zip(fes, bows)]
    try:
        starts = np.concatenate([[0], np.cumsum(shards)])
        for r, (f, d) in enumerate(zip(fes, dbs)):                              # pass 1: every rank fills its own store
            for s in range(0, shards[r], B):
                n = min(B, shards[r] - s)
                f.process_batch_host(frames[starts[r] + s:starts[r] + s + n]); d.insert_batch(n)
            f.synchronize()
        rows = [d.store_pointers() for d in dbs]
        for r, d in enumerate(dbs):
            d.set_peers(r, starts, rows)
        multi = []
        for r, (f, d) in enumerate(zip(fes, dbs)):                              # pass 2: query against the whole logical database
            for s in range(0, shards[r], B):
                n = min(B, shards[r] - s)
                d.query_batch(s, n); res = d.results(n)
                for i in range(n):
                    multi.append({k: v[i] for k, v in res.items()})
        assert len(multi) == NF
        cross = 0
        for g in range(NF):
            _check_frame(multi[g], exp[g], ("multi", g))
            for k in ("query_ids", "query_scores", "query_n", "pair_kf", "matched_n"):
                assert np.array_equal(multi[g][k], single[g][k]), (g, k)
            n = multi[g]["matched_n"]
            assert np.array_equal(multi[g]["best"][:, :n], single[g]["best"][:, :n]), g
            owner = int(np.searchsorted(starts, g, side="right") - 1)
            cross += sum(1 for k in multi[g]["pair_kf"] if k >= 0 and not (starts[owner] <= k < starts[owner + 1]))
        assert cross >= 6, cross                          # keyframes of OTHER ranks really were hit and matched
    finally:
        for d in dbs: d.close()
        for b in bows: b.close()
        for f in fes: f.close()


def test_full_frontend_at_benchmark_size(voc, scene_frames):
    """The full front-end at BASELINE.json's configs[2] shape -- 1241x376, 1000 ORB, 8 levels -- over 64 frames in batches of 16,
    every frame against the sequential oracle (keypoint counts, BoW top-4 with FP64 scores, pair list, every brute-force key)."""
    from srba_stereo_slam_b200 import frontend, synth
    W, H, NL, NFEAT, NF, B = 1241, 376, 8, 1000, 64, 16
    base = scene_frames(8)                                 # 8 rendered frames ...
    frames = synth.stereo_batch(NF, W, H, seed=0, unique=8)   # ... and their 8 mirrored / flipped / inverted variants = 64 distinct valid pairs
    assert np.array_equal(frames[0], base[0])
    exp = _oracle_frontend(voc, frames, NFEAT, NL)
    fe = frontend.StereoFrontEnd(width=W, height=H, nlevels=NL, nfeatures=NFEAT, max_pairs=B)
    bow = frontend.BowDatabase(fe, voc); db = frontend.KeyframeDatabase(fe, bow, capacity_frames=NF)
    try:
        g = 0
        for s in range(0, NF, B):
            fe.process_batch_host(frames[s:s + B]); db.process_batch(B)
            r = db.results(B)
            for i in range(B):
                _check_frame({k: v[i] for k, v in r.items()}, exp[g], g); g += 1
        assert sum(len(e["res"]) == 4 for e in exp) > 50 and sum(e["n"] for e in exp) > 64 * 250
    finally:
        db.close(); bow.close(); fe.close()
