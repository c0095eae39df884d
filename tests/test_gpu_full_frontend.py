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
