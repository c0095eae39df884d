"""Frame / keyframe sharding across the GPUs of one box.

The front-end shards by frame: extraction and L/R matching of a stereo pair depend on nothing else, so rank r
of G takes the contiguous chunk [r*F/G, (r+1)*F/G) (plus a one-frame halo for inter-frame matching, recomputed
locally, no communication).  Keyframe-to-keyframe matching and BoW queries shard by QUERY keyframe against a
replicated database; the only collective on the path is one all-gather of fixed-stride keyframe descriptor
blocks {u32 n; u8 desc[Nmax][32]} (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous, balanced [start, end): the first n_items % world ranks get one extra item."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_with_halo(n_frames, rank, world):
    """Frames a rank must PROCESS to own [start, end) when inter-frame matching needs the previous frame."""
    start, end = shard_range(n_frames, rank, world)
    return max(start - 1, 0), start, end


def pack_kf_blocks(desc_list, nmax):
    """[(n_i, 32) uint8] -> (counts int32 [K], blocks uint8 [K, nmax, 32]) fixed-stride all-gather payload."""
    k = len(desc_list)
    counts = np.zeros(k, np.int32)
    blocks = np.zeros((k, nmax, 32), np.uint8)
    for i, d in enumerate(desc_list):
        d = np.asarray(d, np.uint8).reshape(-1, 32)
        if len(d) > nmax:
            raise ValueError("keyframe %d has %d descriptors > nmax %d" % (i, len(d), nmax))
        counts[i] = len(d); blocks[i, :len(d)] = d
    return counts, blocks


def allgather_kf_blocks(counts, blocks, n_total, dist, device=None):
    """All-gather per-rank keyframe blocks into the global database, global order = rank-major shard order.

    counts/blocks: this rank's shard (numpy, or torch tensors already on `device`).  Shards may be ragged (they are
    padded to the largest shard for the collective and trimmed afterwards).  Returns torch tensors
    (counts [n_total], blocks [n_total, nmax, 32]) on `device`."""
    import torch
    world = dist.get_world_size()
    dev = device if device is not None else "cpu"
    counts_t = torch.as_tensor(counts, dtype=torch.int32, device=dev)
    blocks_t = torch.as_tensor(blocks, dtype=torch.uint8, device=dev)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    kmax = max(e - s for s, e in sizes)
    nmax = blocks_t.shape[1]
    pad_c = torch.zeros(kmax, dtype=torch.int32, device=dev); pad_c[:counts_t.shape[0]] = counts_t
    pad_b = torch.zeros((kmax, nmax, 32), dtype=torch.uint8, device=dev); pad_b[:blocks_t.shape[0]] = blocks_t
    all_c = torch.empty((world, kmax), dtype=torch.int32, device=dev)
    all_b = torch.empty((world, kmax, nmax, 32), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(all_c.view(-1), pad_c)
    dist.all_gather_into_tensor(all_b.view(-1), pad_b.view(-1))
    cs = [all_c[r, :e - s] for r, (s, e) in enumerate(sizes)]
    bs = [all_b[r, :e - s] for r, (s, e) in enumerate(sizes)]
    return torch.cat(cs), torch.cat(bs)
