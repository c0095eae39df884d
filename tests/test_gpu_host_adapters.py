"""GPU (-m gpu): the C++ adapter classes (same names/signatures as the reference: rso::CStereoOdometryEstimator,
CStereoSLAMKF, CBoWManager + the BFMatcher part of m_internal_data_association) driven like performStereoSLAM
drives them (EST.cpp:82-216, 413-447), checked against the oracle."""
import os
import struct
import subprocess
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "srba-stereo-slam_b200", "host")


def _write_voc_yml(voc, path):
    """Re-serialise the flattened vocabulary fixture in DBoW2's YAML layout so the C++ loader is exercised too."""
    with open(path, "w") as f:
        f.write("%%YAML:1.0\nvocabulary:\n   k: %d\n   L: %d\n   scoringType: 0\n   weightingType: 0\n   nodes:\n" % (voc["k"], voc["L"]))
        parent = voc["parent"]
        order = []                                      # file order must list a node's children in child_idx order
        for p in range(voc["nnodes"]):
            order.extend(voc["child_idx"][voc["child_ofs"][p]:voc["child_ofs"][p + 1]].tolist())
        for nid in order:
            bits = np.unpackbits(voc["node_desc"][nid], bitorder="little")[::-1]       # char c <-> bit 255-c
            f.write('      - { nodeId:%d, parentId:%d, weight:%s,\n          descriptor:"%s" }\n' %
                    (nid, parent[nid], repr(float(voc["node_weight"][nid])), "".join(map(str, bits))))
        f.write("   words:\n")
        for nid in np.nonzero(voc["node_word"] >= 0)[0]:
            f.write("      - { wordId:%d, nodeId:%d }\n" % (voc["node_word"][nid], nid))


def test_adapters_follow_the_reference_call_sequence(tmp_path, voc):
    from srba_stereo_slam_b200 import frontend, synth
    from oracle import oracle as orc
    W, H, NF, NL, NFEAT = 640, 192, 3, 4, 300
    frames = synth.CorridorScene(seed=6, width=W, height=H).sequence(NF)
    fb = tmp_path / "frames.bin"; frames.tofile(fb)
    vy = tmp_path / "voc.yml"; _write_voc_yml(voc, vy)
    ob = tmp_path / "out.bin"
    r = subprocess.run([os.path.join(HOST, "host_test"), str(fb), str(W), str(H), str(NF), str(NL), str(NFEAT), str(vy), str(ob)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    data = open(ob, "rb").read(); off = 0
    KD, MD = frontend.KEYPOINT_DTYPE, frontend.DMATCH_DTYPE
    prev = None; odb = []; next_id = 0
    for k in range(NF):
        nl, nr, nm, tracked, nq, nda = struct.unpack_from("6i", data, off); off += 24
        kl = np.frombuffer(data, KD, nl, off); off += 28 * nl
        kr = np.frombuffer(data, KD, nr, off); off += 28 * nr
        dl = np.frombuffer(data, np.uint8, nl * 32, off).reshape(nl, 32); off += 32 * nl
        dr = np.frombuffer(data, np.uint8, nr * 32, off).reshape(nr, 32); off += 32 * nr
        m = np.frombuffer(data, MD, nm, off); off += 16 * nm
        ids = np.frombuffer(data, np.int64, nm, off); off += 8 * nm
        qr = [struct.unpack_from("=id", data, off + 12 * i) for i in range(nq)]; off += 12 * nq
        da = np.frombuffer(data, MD, nda, off); off += 16 * nda
        # --- oracle
        ol, odl, _ = orc.orb_extract(frames[k, 0], nfeatures=NFEAT, nlevels=NL)
        orr, odr, _ = orc.orb_extract(frames[k, 1], nfeatures=NFEAT, nlevels=NL)
        om = orc.match_rowband(ol, odl, orr, odr)
        for f in ("x", "y", "size", "angle", "response", "octave", "class_id"):
            assert np.array_equal(kl[f], ol[f]) and np.array_equal(kr[f], orr[f]), (k, f)
        assert np.array_equal(dl, odl) and np.array_equal(dr, odr)
        assert np.array_equal(m["queryIdx"], om["queryIdx"]) and np.array_equal(m["trainIdx"], om["trainIdx"]) and nm > 30
        cur = odl[om["queryIdx"]]
        # id propagation (ifmDescBF): best previous match under the ORB threshold, each previous id claimed once
        exp_ids = np.zeros(nm, np.int64); exp_tracked = 0
        if prev is not None:
            bf = orc.match_bf(cur, prev[0]); claimed = set()
            for i in range(nm):
                t = int(bf["trainIdx"][i])
                if bf["distance"][i] <= 60 and t not in claimed:
                    exp_ids[i] = prev[1][t]; claimed.add(t); exp_tracked += 1
                else:
                    exp_ids[i] = next_id; next_id += 1
            oda = orc.match_bf(cur, prev[0])                       # KF-to-KF brute force on the matched subsets (EST.cpp:1432-1452)
            assert np.array_equal(da["trainIdx"], oda["trainIdx"]) and np.array_equal(da["distance"], oda["distance"])
            assert np.array_equal(da["queryIdx"], np.arange(nm))
        else:
            exp_ids = np.arange(nm, dtype=np.int64); next_id = nm
        assert np.array_equal(ids, exp_ids) and tracked == exp_tracked
        if k > 0:
            assert tracked > 10
            ores = orc.bow_query(voc, odb, odl, 4)
            assert [q[0] for q in qr] == [o[0] for o in ores] and [q[1] for q in qr] == [o[1] for o in ores]
        w_, wt_ = orc.bow_transform(voc, odl)
        odb.append(orc.bow_vector(voc, w_, wt_))
        prev = (cur, exp_ids)
    assert off == len(data)


def test_adaptive_threshold_loop(tmp_path):
    """The reference's adaptive FAST/ORB threshold loop (EST.cpp:263-315) driven through the C++ adapters on low-contrast
    frames, against the same loop replayed in Python on top of the oracle."""
    from srba_stereo_slam_b200 import frontend, synth
    from oracle import oracle as orc
    W, H, NF, NL, NFEAT, MINM = 640, 192, 4, 4, 300, 100
    frames = synth.CorridorScene(seed=6, width=W, height=H).sequence(NF)
    frames[1] = (frames[1].astype(np.float32) * 0.35 + 80).astype(np.uint8)     # low contrast: few corners at FAST 20
    frames[2] = (frames[2].astype(np.float32) * 0.2 + 100).astype(np.uint8)
    fb = tmp_path / "frames.bin"; frames.tofile(fb)
    ob = tmp_path / "out.bin"
    r = subprocess.run([os.path.join(HOST, "host_test"), str(fb), str(W), str(H), str(NF), str(NL), str(NFEAT), "-", str(ob), str(MINM)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    data = open(ob, "rb").read(); off = 0
    fast_th, orb_th = 20, 60                          # adapter defaults: fast 5..30, orb 30..100 (example.ini:15-16,45-46)
    seen_repeat = False
    MD = frontend.DMATCH_DTYPE
    prev = None; next_id = 0
    for k in range(NF):
        calls, f_th, o_th = struct.unpack_from("3i", data, off); off += 12
        nl, nr, nm, tracked, nq, nda = struct.unpack_from("6i", data, off); off += 24
        off += 28 * nl + 28 * nr + 32 * nl + 32 * nr
        mm = np.frombuffer(data, MD, nm, off); off += 16 * nm
        ids = np.frombuffer(data, np.int64, nm, off); off += 8 * nm
        off += 12 * nq + 16 * nda
        # ---- replay
        ecalls = 0
        while True:
            ol, odl, _ = orc.orb_extract(frames[k, 0], nfeatures=NFEAT, nlevels=NL, fast_threshold=fast_th)
            orr, odr, _ = orc.orb_extract(frames[k, 1], nfeatures=NFEAT, nlevels=NL, fast_threshold=fast_th)
            om = orc.match_rowband(ol, odl, orr, odr, max_dist=orb_th)
            m = len(om); ecalls += 1
            repeat = False
            if m < MINM:
                if fast_th > 5: fast_th = max(5, min(30, max(0, fast_th - 10))); repeat = True
                elif orb_th < 100: orb_th = min(100, orb_th + 10); repeat = True
                elif m >= 8: pass
                else: raise AssertionError("would abort")
            elif m < MINM * 1.2:
                if fast_th > 5: fast_th = max(5, fast_th - 5)
                elif orb_th < 100: orb_th = min(100, orb_th + 5)
            else:
                fast_th = min(20, fast_th + 5); orb_th = 60
            if not repeat:
                break
        seen_repeat |= ecalls > 1
        assert (calls, f_th, o_th, nm) == (ecalls, fast_th, orb_th, m), k
        assert np.array_equal(mm["queryIdx"], om["queryIdx"]) and np.array_equal(mm["trainIdx"], om["trainIdx"])
        # id propagation across a REPEATED frame (EST.cpp:263-315, :359-378): only the last attempt counts, it is matched
        # against the true previous frame (not against the frame's own discarded first attempt) and the ids the discarded
        # attempts consumed are handed out again
        cur = odl[om["queryIdx"]]
        exp_ids = np.zeros(m, np.int64); exp_tracked = 0
        if prev is None:
            exp_ids = np.arange(m, dtype=np.int64); next_id = m
        else:
            bf = orc.match_bf(cur, prev[0]); claimed = set()
            for i in range(m):
                t = int(bf["trainIdx"][i]) if len(bf) else -1
                if len(bf) and bf["distance"][i] <= 60 and t not in claimed:
                    exp_ids[i] = prev[1][t]; claimed.add(t); exp_tracked += 1
                else:
                    exp_ids[i] = next_id; next_id += 1
        assert np.array_equal(ids, exp_ids), (k, ecalls)
        assert tracked == exp_tracked, (k, ecalls, tracked, exp_tracked)
        if ecalls > 1 and prev is not None:
            assert tracked < m            # a frame matched against its own first attempt would track (almost) everything
        prev = (cur, exp_ids)
    assert off == len(data) and seen_repeat
