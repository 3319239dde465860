"""GPU: the user-facing FeaturePath calls (host buffers, pipelined device->host copies) return exactly
what the device-resident path computes; per-pair splitting matches the pair segments."""
import numpy as np
import pytest
import torch

from pcrcg_b200 import blocks, pipeline, synthetic

pytestmark = pytest.mark.gpu


def test_run_host_equals_run_device_and_pipelining():
    cfg = blocks.indoor_config(first_feats_dim=64)
    path = pipeline.FeaturePath(cfg, [30, 30, 30, 30], device="cuda:0")
    pairs = [synthetic.match3d_pair(s, n_target=1500)[:2] for s in range(3)]
    pts, lens = pipeline.stack_pairs(pairs)
    ph, lh = torch.from_numpy(pts).pin_memory(), torch.from_numpy(lens).pin_memory()
    y, batch = path.run_device(ph.cuda(), lh.cuda())
    out, coarse = path.run_host(ph, lh)
    assert torch.equal(out, y.cpu()) and torch.equal(coarse, batch["stack_lengths"][-1].cpu())
    bufs = [torch.empty((y.shape[0] + 7, y.shape[1])).pin_memory() for _ in range(2)]
    hs = [path.submit_host(ph, lh, bufs[i & 1]) for i in range(2)]
    for h in hs:
        o, c = h.result()
        assert torch.equal(o, y.cpu()) and torch.equal(c, coarse)
    per_pair = pipeline.per_pair_features(path, pairs)
    assert sum(f.shape[0] for f in per_pair) == y.shape[0]
    assert torch.equal(torch.cat(per_pair), y)
    # one pair alone == its slice of the stacked run (InstanceNorm statistics are per pair)
    single = pipeline.per_pair_features(path, pairs[1:2])[0]
    assert float((single - per_pair[1]).abs().max() / per_pair[1].abs().max()) < 3e-4
