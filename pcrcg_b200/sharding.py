"""Multi-GPU: the path partitions by fragment pair (``assert len(list_data) == 1``,
datasets/dataloader.py:207; InstanceNorm statistics are per pair), so pairs are sharded across
ranks -- pair i goes to rank ``i % world`` -- with NO collective on the hot path.  torch.distributed
(NCCL over NVLink on GPUs, gloo in the CPU tests) is used only after the loop, to gather the
per-pair results and the timing vectors on rank 0.
"""
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world):
    """Static round-robin partition: the pair indices owned by ``rank``."""
    return list(range(rank, n_items, world))


def _dev_of(group_backend, fallback):
    return fallback if group_backend == "nccl" else torch.device("cpu")


def gather_results(local, feat_dim, device, group=None):
    """local: list of (pair_index, tensor [rows_i, feat_dim]).  Returns {pair_index: tensor} with every
    rank's items on every rank (rank 0 is the consumer; all_gather keeps the protocol symmetric)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return {i: t for i, t in local}
    world = dist.get_world_size(group)
    dev = _dev_of(dist.get_backend(group), device)
    n_local = torch.tensor([len(local)], dtype=torch.int64, device=dev)
    rows_local = torch.tensor([sum(int(t.shape[0]) for _, t in local)], dtype=torch.int64, device=dev)
    mx = torch.stack([n_local, rows_local]).view(-1)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    max_items, max_rows = int(mx[0]), int(mx[1])
    meta = torch.full((max_items, 2), -1, dtype=torch.int64, device=dev)
    for k, (i, t) in enumerate(local):
        meta[k, 0], meta[k, 1] = i, int(t.shape[0])
    flat = torch.zeros((max_rows, feat_dim), dtype=torch.float32, device=dev)
    if local:
        cat = torch.cat([t.to(dev, torch.float32) for _, t in local], 0)
        flat[:cat.shape[0]] = cat
    metas = [torch.empty_like(meta) for _ in range(world)]
    flats = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    dist.all_gather(flats, flat, group=group)
    out = {}
    for m, f in zip(metas, flats):
        o = 0
        for i, r in m.tolist():
            if i < 0:
                continue
            out[i] = f[o:o + r]
            o += r
    return out


def gather_timings(ms_local, device, group=None):
    """Per-rank timing vector -> [world, len] on every rank (max over ranks is the job's time)."""
    t = torch.as_tensor(ms_local, dtype=torch.float64).view(1, -1)
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    dev = _dev_of(dist.get_backend(group), device)
    t = t.to(dev)
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(outs, t, group=group)
    return torch.cat(outs, 0)


def run_sharded(pairs, compute, feat_dim, device, pairs_per_batch=8, group=None):
    """pairs: list of (src, tgt) NumPy clouds, identical on every rank.  compute(list_of_pairs) ->
    list of per-pair feature tensors.  Each rank processes its shard in mini-batches, then the
    results are gathered.  Returns {pair_index: features}."""
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    mine = shard_indices(len(pairs), rank, world)
    local = []
    for b in range(0, len(mine), pairs_per_batch):
        idx = mine[b:b + pairs_per_batch]
        feats = compute([pairs[i] for i in idx])
        local += list(zip(idx, feats))
    return gather_results(local, feat_dim, device, group)
