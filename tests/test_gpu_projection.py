"""GPU parity: colour path (projection.py + feature scatter) vs the reference's golden vectors and
the oracle on seeded synthetic RGB-D views.  Index lists are bit-exact; scattered rows are exact
(they are copies and one fp32 product)."""
import os

import numpy as np
import pytest
import torch

from oracle import projection_port as pp
from pcrcg_b200 import projection as gp
from pcrcg_b200 import synthetic, blocks, ops, dataloader

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = os.path.join(os.path.dirname(__file__), "golden")


def test_projection_vs_reference_golden():
    g = np.load(os.path.join(G, "projection_ref.npz"))
    pts = torch.from_numpy(g["points"])
    views = []
    for vi in (0, 1):
        pr = gp.Projection(torch.from_numpy(g[f"v{vi}_intrinsics"]))
        i2, i3 = pr.projection(pts, torch.from_numpy(g[f"v{vi}_depth"])[None], torch.from_numpy(g[f"v{vi}_world2camera"]))
        assert i2.dtype == torch.int64 and i3.dtype == torch.int64 and i2.device.type == "cpu"
        assert np.array_equal(i2.numpy(), g[f"v{vi}_inds2d"]) and np.array_equal(i3.numpy(), g[f"v{vi}_inds3d"])
        views.append(dict(depth=g[f"v{vi}_depth"], world2camera=g[f"v{vi}_world2camera"], intrinsics=g[f"v{vi}_intrinsics"],
                          feature2d=torch.from_numpy(g[f"v{vi}_feature2d"]).to(DEV), valid_map=g[f"v{vi}_valid_map"]))
    x = gp.unproject_features(pts.to(DEV), [views[1], views[0]])          # image 2 written first, image 1 wins
    assert np.array_equal(x.cpu().numpy(), g["x_out"])


@pytest.mark.parametrize("seed", [0, 1])
def test_projection_pair_vs_oracle(seed):
    src, tgt, _ = synthetic.match3d_pair(seed, n_target=8000)
    pts = np.concatenate([src, tgt])
    n_src = len(src)
    write_order, oracle_views = [], []
    for cloud, (lo, hi), s in ((src, (0, n_src), 10 + seed), (tgt, (n_src, len(pts)), 20 + seed)):
        vs = synthetic.rgbd_views(cloud, s, n_views=2, channels=32)
        for v in (vs[1], vs[0]):                                          # image 2 then image 1
            i2, i3 = pp.projection(cloud, v["depth"], v["world2camera"], v["intrinsics"])
            g2, g3 = gp.Projection(torch.from_numpy(v["intrinsics"])).projection(torch.from_numpy(cloud).to(DEV),
                                                                                torch.from_numpy(v["depth"]).to(DEV), torch.from_numpy(v["world2camera"]))
            assert g2.device.type == "cuda"
            assert np.array_equal(g2.cpu().numpy(), i2) and np.array_equal(g3.cpu().numpy(), i3)
            assert len(i3) > 100
            oracle_views.append((v["feature2d"], v["valid_map"], i2, i3 + lo))
            write_order.append(dict(v, feature2d=torch.from_numpy(v["feature2d"]).to(DEV), rows=(lo, hi)))
    ref = pp.scatter_image_features(len(pts), oracle_views)
    x = gp.unproject_features(torch.from_numpy(pts).to(DEV), write_order)
    assert np.array_equal(x.cpu().numpy(), ref)


def test_colour_path_first_block_129_channels():
    """PCR-CG colour config: in_feats_dim = 129 (configs/test/indoor.yaml:36) through the first KPConv block."""
    from oracle import blocks_port as bp
    src, tgt, _ = synthetic.match3d_pair(2, n_target=2500)
    pts = np.concatenate([src, tgt]); lens = np.array([len(src), len(tgt)], np.int32)
    views = []
    for cloud, (lo, hi), s in ((src, (0, len(src)), 1), (tgt, (len(src), len(pts)), 2)):
        for v in synthetic.rgbd_views(cloud, s, n_views=2, channels=128)[::-1]:
            views.append(dict(v, feature2d=torch.from_numpy(v["feature2d"]).to(DEV), rows=(lo, hi)))
    x = gp.unproject_features(torch.from_numpy(pts).to(DEV), views)
    assert x.shape == (len(pts), 129)
    cfg = blocks.indoor_config(in_feats_dim=129, first_feats_dim=128)
    b = dataloader.build_pyramid(pts, lens, cfg, [34, 39, 39, 38], device=DEV)
    torch.manual_seed(0)
    blk = blocks.SimpleBlock("simple", 129, 128, 0.0625, 0, cfg).to(DEV)
    blk.KPConv.set_kernel_points(torch.randn(15, 3) * 0.03)
    y = blk(x, b)
    ref = bp.simple_block(x.cpu(), b["points"][0].cpu(), b["points"][0].cpu(), b["neighbors"][0].cpu(),
                          dict(kernel_points=blk.KPConv.kernel_points.cpu(), weights=blk.KPConv.weights.cpu(), KP_extent=0.05))
    err = float((y.cpu() - ref).abs().max() / ref.abs().max())
    assert err < 1e-3


def test_unproject_validates_views_and_accepts_reference_valid_map_layout():
    """every view must share one (C, H, W) (the kernel indexes all of them with it); the reference batch's [W,H] valid maps
    (transposed inside KPFCNN.forward, models/architectures.py:287-307) are accepted as they are"""
    src, _, _ = synthetic.match3d_pair(3, n_target=3000)
    vs = synthetic.rgbd_views(src, 5, n_views=2, channels=16)
    pts = torch.from_numpy(src).to(DEV)
    mk = lambda v, **o: dict(v, feature2d=torch.from_numpy(v["feature2d"]).to(DEV), **o)
    ref = gp.unproject_features(pts, [mk(vs[1]), mk(vs[0])])
    wh = gp.unproject_features(pts, [mk(vs[1], valid_map=np.ascontiguousarray(vs[1]["valid_map"].T)),
                                     mk(vs[0], valid_map=np.ascontiguousarray(vs[0]["valid_map"].T))])
    assert torch.equal(ref, wh)
    bad_feat = mk(vs[0]); bad_feat["feature2d"] = bad_feat["feature2d"][:, :-1]
    with pytest.raises(RuntimeError, match="differs from view 0"):
        gp.unproject_features(pts, [mk(vs[1]), bad_feat])
    with pytest.raises(RuntimeError, match="depth"):
        gp.unproject_features(pts, [mk(vs[1], depth=vs[1]["depth"][:-2])])
    with pytest.raises(RuntimeError, match="valid_map"):
        gp.unproject_features(pts, [mk(vs[1], valid_map=vs[1]["valid_map"][:, :-3])])
    with pytest.raises(RuntimeError, match="views per call"):
        gp.unproject_features(pts, [mk(vs[0])] * 9)
    with pytest.raises(RuntimeError, match="rows"):
        gp.unproject_features(pts, [mk(vs[0], rows=(0, len(src) + 1))])


def test_unproject_batch_equals_per_pair_calls_and_oracle():
    """one launch for a stacked batch (3 pairs, 12 views, device-resident view table) == the per-pair 8-view entry point ==
    the oracle's scatter, bit for bit"""
    pairs = [synthetic.match3d_pair(20 + k, n_target=2500 + 300 * k)[:2] for k in range(3)]
    clouds = [c for p in pairs for c in p]
    pts = torch.from_numpy(np.concatenate(clouds)).to(DEV)
    lens = np.array([len(c) for c in clouds], np.int32)
    per_cloud, oracle_rows, lo = [], [], 0
    for ci, cloud in enumerate(clouds):
        vs = synthetic.rgbd_views(cloud, 40 + ci, n_views=2, channels=24)[::-1]
        per_cloud.append([dict(v, feature2d=torch.from_numpy(v["feature2d"]).to(DEV)) for v in vs])
        ov = []
        for v in vs:
            i2, i3 = pp.projection(cloud, v["depth"], v["world2camera"], v["intrinsics"])
            ov.append((v["feature2d"], v["valid_map"], i2, i3))
        oracle_rows.append(pp.scatter_image_features(len(cloud), ov))
        lo += len(cloud)
    x = gp.unproject_features_batch(pts, lens, per_cloud)
    assert np.array_equal(x.cpu().numpy(), np.concatenate(oracle_rows))
    starts = np.concatenate([[0], np.cumsum(lens)])
    for k in range(3):                                                    # the 8-view entry point, pair by pair
        a, b, c = int(starts[2 * k]), int(starts[2 * k + 1]), int(starts[2 * k + 2])
        views = [dict(v, rows=(0, b - a)) for v in per_cloud[2 * k]] + [dict(v, rows=(b - a, c - a)) for v in per_cloud[2 * k + 1]]
        assert torch.equal(gp.unproject_features(pts[a:c], views), x[a:c])
