"""GPU parity on the other BASELINE.json configs at sizes the oracle finishes in seconds: 3DLoMatch-shaped (low overlap)
and KITTI-shaped (voxel 0.3, conv_radius 4.25) pairs.  Pyramid (subsampled points, stack lengths, all three index lists of
every level) bit-exact vs the C port of the reference core; encoder output within 1e-3 normwise vs the PyTorch-fp32 port."""
import numpy as np
import pytest
import torch

import oracle
from oracle import blocks_port as bp
from pcrcg_b200 import blocks, dataloader, synthetic

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
DEV = "cuda:0"


def _port_pyramid(P, pts, lens, cfg, limits):
    """datasets/dataloader.py:239-359 through the C port (canonical (d2, index) ties, width = min(limit, max_count))"""
    r = cfg.first_subsampling_dl * cfg.conv_radius
    out = dict(points=[], neighbors=[], pools=[], upsamples=[], stack_lengths=[])
    for layer in range(cfg.num_layers):
        conv = P.batch_query(pts, pts, lens, lens, r, limit=limits[layer])
        if layer < cfg.num_layers - 1:
            pp, pl = P.subsample_batch(pts, lens, 2 * r / cfg.conv_radius)
            pool = P.batch_query(pp, pts, pl, lens, r, limit=limits[layer])
            up = P.batch_query(pts, pp, lens, pl, 2 * r, limit=limits[layer])
        else:
            pp, pl = np.zeros((0, 3), np.float32), np.zeros((0,), np.int32)
            pool = up = np.zeros((0, 1), np.int32)
        out["points"].append(pts); out["neighbors"].append(conv); out["pools"].append(pool)
        out["upsamples"].append(up); out["stack_lengths"].append(lens)
        pts, lens, r = pp, pl, 2 * r
    return out


def _check(cfg, limits, src, tgt, feats_dim):
    P = oracle.port()
    pts = np.concatenate([src, tgt]).astype(np.float32)
    lens = np.array([len(src), len(tgt)], np.int32)
    ref = _port_pyramid(P, pts, lens, cfg, limits)
    got = dataloader.build_pyramid(pts, lens, cfg, limits, device=DEV)
    for l in range(cfg.num_layers):
        assert np.array_equal(got["points"][l].cpu().numpy(), ref["points"][l]), f"points level {l}"
        assert np.array_equal(got["stack_lengths"][l].cpu().numpy(), ref["stack_lengths"][l])
        for k in ("neighbors", "pools", "upsamples"):
            assert np.array_equal(got[k][l].cpu().numpy(), ref[k][l]), f"{k} level {l}"
    # encoder on this pyramid vs the fp32 port (same weights)
    torch.manual_seed(0)
    net = blocks.KPEncoder(cfg).to(DEV)
    for m in net.modules():
        if isinstance(m, blocks.KPConv):
            m.set_kernel_points(torch.randn(15, 3) * 0.4 * m.radius)
    x = net(torch.ones(len(pts), 1, device=DEV), got)
    sd = {k[len("encoder_blocks."):]: v.cpu() for k, v in net.state_dict().items()}
    pb = bp.encoder_blocks_from_state_dict(sd, first_subsampling_dl=cfg.first_subsampling_dl, conv_radius=cfg.conv_radius,
                                           KP_extent=cfg.KP_extent)
    tb = {k: [torch.from_numpy(np.ascontiguousarray(a)).long() if k != "points" else torch.from_numpy(a) for a in ref[k]]
          for k in ("points", "neighbors", "pools", "upsamples")}
    xr, _ = bp.encoder(torch.ones(len(pts), 1), tb, pb)
    err = float((x.cpu() - xr).abs().max() / xr.abs().max())
    assert x.shape[1] == feats_dim * 8 and err < 1e-3, err


def test_lomatch_shaped_pair():
    src, tgt, _ = synthetic.match3d_pair(3, n_target=2500, overlap="low")
    _check(blocks.indoor_config(first_feats_dim=32), [30, 28, 28, 30], src, tgt, 32)


def test_kitti_shaped_pair():
    """voxel 0.3 / conv_radius 4.25 (configs/test/kitti.yaml:15-17); a cropped synthetic street scan pair"""
    a, b, _ = synthetic.kitti_pair(1)
    crop = lambda s: s[(np.abs(s[:, 0]) < 18) & (np.abs(s[:, 1]) < 12)]
    src = synthetic.voxel_downsample_np(crop(a).astype(np.float64), 0.3).astype(np.float32)
    tgt = synthetic.voxel_downsample_np(crop(b).astype(np.float64), 0.3).astype(np.float32)
    assert 500 < len(src) < 12000
    _check(blocks.kitti_config(first_feats_dim=32), [40, 40, 40, 38], src, tgt, 32)
