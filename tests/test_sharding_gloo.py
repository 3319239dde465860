"""CPU, world_size 2, gloo: the N>1 host-side logic -- pair sharding, result / timing gathers.
The compute is a deterministic CPU stand-in (the CUDA path needs a GPU); what is checked is that the
sharded job returns exactly what a single rank returns, for ragged shards and ragged row counts."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_compute(pairs):
    # "features": a deterministic function of the points, [rows = len(src)//50 + len(tgt)//50, 8]
    out = []
    for s, t in pairs:
        n = len(s) // 50 + len(t) // 50
        base = torch.from_numpy(np.concatenate([s, t])[:n].astype(np.float32))
        out.append(torch.cat([base, base * 2, base[:, :2] + 1], 1))
    return out


def _pairs(n):
    rng = np.random.default_rng(0)
    return [(rng.random((int(rng.integers(200, 900)), 3)).astype(np.float32),
             rng.random((int(rng.integers(200, 900)), 3)).astype(np.float32)) for _ in range(n)]


def _worker(rank, world, port, n_pairs, q):
    sys.path.insert(0, ROOT)
    from pcrcg_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pairs = _pairs(n_pairs)
    res = sharding.run_sharded(pairs, _fake_compute, 8, torch.device("cpu"), pairs_per_batch=2)
    tim = sharding.gather_timings([float(rank + 1), 10.0 * (rank + 1)], torch.device("cpu"))
    if rank == 0:
        q.put(({k: v.numpy() for k, v in res.items()}, tim.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [5, 1, 8])
def test_sharded_equals_single(n_pairs):
    sys.path.insert(0, ROOT)
    from pcrcg_b200 import sharding
    assert sharding.shard_indices(5, 1, 2) == [1, 3] and sharding.shard_indices(5, 0, 2) == [0, 2, 4]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_pairs) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    res, tim = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = _fake_compute(_pairs(n_pairs))
    assert sorted(res.keys()) == list(range(n_pairs))
    for i in range(n_pairs):
        assert np.array_equal(res[i], single[i].numpy())
    assert tim.shape == (2, 2) and tim.max(0).tolist() == [2.0, 20.0]
