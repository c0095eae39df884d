"""CPU: host-side logic -- vocabulary loader, sharding arithmetic, synthetic scene determinism, and the N>1
all-gather path on world_size-2 gloo."""
import os
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from srba_stereo_slam_b200 import sharding, synth, voc_loader
from oracle import voc as oracle_voc

REF_VOC = "/root/reference/demo/voc.yml.gz"


def test_voc_flat_fixture_shape(voc):
    assert voc["k"] == 8 and voc["L"] == 5 and voc["scoring"] == 0 and voc["weighting"] == 0
    assert voc["nnodes"] == 36940 and voc["nwords"] == 32273
    nch = np.diff(voc["child_ofs"])
    assert (nch > 0).sum() == 4667 and (nch == 8).sum() == 4552 and nch.max() == 8
    assert ((nch == 0) == (voc["node_word"] >= 0)).all()
    assert (voc["node_weight"][nch > 0] == 0).all() and (voc["node_weight"][nch == 0] == 0).sum() == 4


@pytest.mark.skipif(not os.path.exists(REF_VOC), reason="reference vocabulary only exists in the build container")
def test_product_loader_matches_independent_parser(voc):
    a = voc_loader.load_vocabulary(REF_VOC)
    b = oracle_voc.load(REF_VOC)
    for k in b:
        assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a[k], voc[k]), k


def test_shard_ranges():
    for n in (0, 1, 7, 4541, 4096, 20000):
        for world in (1, 2, 4, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(e - s for s, e in r) - min(e - s for s, e in r) <= 1
    assert sharding.shard_with_halo(4541, 3, 8) == (1703, 1704, 2272)      # 4541 = 8*567 + 5
    assert sharding.shard_with_halo(4541, 0, 8)[0] == 0


def test_scene_is_deterministic():
    a = synth.CorridorScene(seed=9, width=320, height=96).frame(3)
    b = synth.CorridorScene(seed=9, width=320, height=96).frame(3)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and not np.array_equal(a[0], a[1])
    assert a[0].dtype == np.uint8 and a[0].shape == (96, 320) and a[0].std() > 10


def _gloo_worker(rank, world, port, n_total, nmax, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)                      # every rank can regenerate the global database
    ns = rng.integers(1, nmax + 1, n_total)
    full = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for n in ns]
    s, e = sharding.shard_range(n_total, rank, world)
    counts, blocks = sharding.pack_kf_blocks(full[s:e], nmax)
    gc, gb = sharding.allgather_kf_blocks(counts, blocks, n_total, dist)
    ok = gc.numpy().tolist() == ns.tolist()
    for i in range(n_total):
        ok = ok and np.array_equal(gb[i, :ns[i]].numpy(), full[i]) and not gb[i, ns[i]:].any()
    np.save(os.path.join(out_dir, "ok_%d.npy" % rank), np.array([int(ok)]))
    dist.destroy_process_group()


def test_allgather_kf_blocks_world2_gloo(tmp_path):
    """The N>1 exchange step on CPU: ragged shards (7 keyframes over 2 ranks) gathered into the global database."""
    port = 29500 + os.getpid() % 2000
    mp.spawn(_gloo_worker, args=(2, port, 7, 20, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert np.load(tmp_path / ("ok_%d.npy" % r))[0] == 1
