"""GPU parity AT BENCH SIZE: the exact batches bench.py times (same generator, same seeds, same limits) are built as one
stacked pyramid on the GPU and compared, pair by pair, with the reference run one pair at a time on the CPU
(oracle/checks.py: unmodified reference C++ when oracle/_ref is built, canonical (d2, index) ties):

  * configs[1]  3DMatch-shaped, the 32-pair / 1.25 M-point / 64-cloud step of the default bench line: every index list of
                every level of every pair (10 lists + 4 point sets, x 32 pairs) array_equal
  * configs[3]  KITTI-shaped with the calibrated limits [102, 102, 99, 91] (lists wider than 64: the H > 64 kernels)
  * configs[2]  3DLoMatch-shaped at first_feats_dim = 256
  and the 11-block encoder (first_feats_dim = 256) within 1e-3 (normwise) of oracle/blocks_port.py on >= 2 pairs of each
  batch, computed from the STACKED run (per-pair InstanceNorm segments)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import checks
from pcrcg_b200 import pipeline

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _stacked(workload, n_pairs):
    cfg, limits = bench.workload_config(workload)
    pairs = bench.make_pairs(workload, n_pairs, 0)                # rank 0's pairs of the bench
    pts, lens = pipeline.stack_pairs(pairs)
    path = pipeline.FeaturePath(cfg, limits, device=DEV, seed=0)
    y, batch = path.run_device(torch.from_numpy(pts).to(DEV), torch.from_numpy(lens).to(DEV))
    torch.cuda.synchronize()
    return cfg, limits, pairs, path, y, batch


def _check(workload, n_pairs, encoder_pairs):
    cfg, limits, pairs, path, y, batch = _stacked(workload, n_pairs)
    seg = batch["pair_segments"][-1].cpu().tolist()
    sd = path.encoder.state_dict()
    total = 0
    for k, (src, tgt) in enumerate(pairs):
        cpu = checks.cpu_pyramid(src, tgt, limits, cfg.first_subsampling_dl, cfg.conv_radius, cfg.num_layers)
        n, bad = checks.compare_pair(batch, k, cpu)
        assert not bad, f"{workload} pair {k}: {bad}"
        total += n
        if k in encoder_pairs:
            err = checks.encoder_error(y[seg[k]:seg[k + 1]], cpu, sd, cfg)
            assert err < 1e-3, f"{workload} pair {k}: encoder error {err}"
    return total, batch


def test_3dmatch_bench_batch_32_pairs():
    total, batch = _check("3dmatch", 32, encoder_pairs=(0, 17))
    assert total == 32 * 14          # 4 point sets + 4 conv + 3 pool + 3 upsample lists per pair
    assert batch["points"][0].shape[0] > 1_000_000 and batch["stack_lengths"][0].numel() == 64


def test_kitti_bench_limits_wide_lists():
    total, batch = _check("kitti", 6, encoder_pairs=(0, 3))
    assert total == 6 * 14
    assert max(int(t.shape[1]) for t in batch["neighbors"]) > 64       # the wide-list (H > 64) kernels are on the path


def test_3dlomatch_first_feats_dim_256():
    cfg, _ = bench.workload_config("3dlomatch")
    assert cfg.first_feats_dim == 256
    total, _ = _check("3dlomatch", 6, encoder_pairs=(1, 4))
    assert total == 6 * 14
