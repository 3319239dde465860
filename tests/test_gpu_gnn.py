"""GPU parity for the 'next' rows 3 of the scope table: the bottleneck overlap-attention GNN (models/gcn.py) and the
whole KPFCNN.forward (models/architectures.py), through the C ABI, vs golden vectors produced by the REFERENCE
(tests/golden/gnn_ref.npz, made by tests/golden/make_golden.py) and vs the oracle port on seeded inputs.
Tolerance: features 1e-3 normwise (north_star); kNN index lists exact."""
import os

import numpy as np
import pytest
import torch

from oracle import gcn_port as gp
from pcrcg_b200 import architectures, blocks, dataloader, gcn, ops, synthetic

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = "cuda:0"
G = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-3


def _d(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def _err(a, ref):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    ref = ref.detach().cpu().numpy() if torch.is_tensor(ref) else np.asarray(ref)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30))


@pytest.fixture(scope="module")
def gnn():
    return np.load(os.path.join(G, "gnn_ref.npz"))


@pytest.fixture(scope="module")
def blk():
    return np.load(os.path.join(G, "blocks_ref.npz"))


def test_knn_equals_reference_lists(gnn):
    idx = ops.knn(_d(gnn["gcn_coords"]), ops.cloud_starts(gnn["gcn_lens"]), 10)
    assert np.array_equal(idx.cpu().numpy(), gnn["gcn_knn"])


@pytest.mark.parametrize("seed", [0, 1])
def test_knn_vs_port_random_clouds(seed):
    rng = np.random.default_rng(seed)
    lens = np.array([300, 11, 1500, 64], np.int32)
    pts = np.concatenate([rng.normal(size=(n, 3)).astype(np.float32) * (1 + i) for i, n in enumerate(lens)])
    pts[5] = pts[4]                                      # duplicate point (distance clamps to 1e-12, tie by index)
    idx = ops.knn(_d(pts), ops.cloud_starts(lens), 10).cpu().numpy()
    ref = gp.knn_lists(torch.from_numpy(pts), lens, 10).numpy()
    same = (idx == ref).all(1)
    # rows may differ only where the reference's topk met an exact distance tie
    for r in np.nonzero(~same)[0]:
        c = np.searchsorted(np.cumsum(lens), r, side="right")
        off = int(np.cumsum(lens)[c] - lens[c])
        d = gp.square_distance(torch.from_numpy(pts[r:r + 1]), torch.from_numpy(pts[off:off + lens[c]]))[0].numpy()
        assert np.allclose(np.sort(d[idx[r] - off]), np.sort(d[ref[r] - off]), rtol=0, atol=0), f"row {r}"
    assert same.mean() > 0.99


def test_small_ops_vs_torch():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(777, 130, generator=g).to(DEV)
    b = torch.randn(130, generator=g).to(DEV)
    assert torch.allclose(ops.bias_act(x, b, 0.0), torch.relu(x + b), atol=1e-6)
    assert torch.allclose(ops.bias_act(x, b), x + b, atol=1e-6)
    y = ops.softmax_rows_(x.clone(), 0.37)
    assert torch.allclose(y, torch.softmax(x * 0.37, dim=1), atol=1e-6)
    assert torch.allclose(ops.l2_normalize(x), torch.nn.functional.normalize(x, p=2, dim=1), atol=1e-6)
    a = torch.randn(200, 512, generator=g).to(DEV)
    w = torch.randn(333, 512, generator=g).to(DEV)
    o = ops.gemm(a[:, 128:256], w[:, 128:256], True)                    # column slices: strides passed through
    assert _err(o, a[:, 128:256].double().cpu() @ w[:, 128:256].double().cpu().t()) < 5e-5


def _load_gcn(gnn, dim=128):
    net = gcn.GCN(4, dim, 10, ["self", "cross", "self"]).to(DEV)
    sd = {k[6:]: torch.from_numpy(gnn[k]) for k in gnn.files if k.startswith("gcnsd_")}
    missing, unexpected = net.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return net


def test_gcn_vs_reference_golden(gnn):
    net = _load_gcn(gnn)
    out = net.forward_rows(_d(gnn["gcn_coords"]), gnn["gcn_lens"], _d(gnn["gcn_feats"]))
    assert _err(out, gnn["gcn_out"]) < TOL
    # reference call convention: [1,3,N] coordinates, [1,C,N] features
    n0 = int(gnn["gcn_lens"][0])
    c, f = _d(gnn["gcn_coords"]), _d(gnn["gcn_feats"])
    d0, d1 = net(c[:n0].t().unsqueeze(0), c[n0:].t().unsqueeze(0), f[:n0].t().unsqueeze(0), f[n0:].t().unsqueeze(0))
    assert d0.shape == (1, 128, n0)
    both = torch.cat([d0, d1], dim=-1)[0].t()
    assert _err(both, gnn["gcn_out"]) < TOL


def test_gcn_stacked_pairs_equal_single_pairs(gnn):
    """two pairs stacked: self-attention statistics stay per cloud, cross-attention per pair"""
    net = _load_gcn(gnn)
    rng = np.random.default_rng(5)
    lens = [140, 97, 60, 201]
    coords = torch.from_numpy(rng.uniform(0, 2, size=(sum(lens), 3)).astype(np.float32)).to(DEV)
    feats = torch.from_numpy(rng.normal(size=(sum(lens), 128)).astype(np.float32)).to(DEV)
    whole = net.forward_rows(coords, lens, feats)
    a = net.forward_rows(coords[:237], lens[:2], feats[:237])
    b = net.forward_rows(coords[237:], lens[2:], feats[237:])
    assert _err(whole, torch.cat([a, b])) < 1e-5
    ref = torch.cat([gp.gcn(coords[s:e].cpu(), l, feats[s:e].cpu(), {k: v.cpu() for k, v in net.state_dict().items()},
                            ["self", "cross", "self"], 4, 10) for s, e, l in ((0, 237, lens[:2]), (237, 498, lens[2:]))])
    assert _err(whole, ref) < TOL


def _geom(blk):
    P = [_d(blk[f"points_{l}"]) for l in range(4)]
    nb = [_d(blk[f"neighbors_{l}"], torch.int32) for l in range(4)]
    pools = [_d(blk[f"pools_{l}"], torch.int32) for l in range(4)]
    ups = [_d(blk[f"upsamples_{l}"], torch.int32) for l in range(4)]
    lens = [torch.from_numpy(blk[f"stack_lengths_{l}"]) for l in range(4)]
    return dict(points=P, neighbors=nb, pools=pools, upsamples=ups, stack_lengths=lens)


def test_whole_network_vs_reference_golden(blk, gnn):
    """KPFCNN.forward of the reference (image_feature False) on the small pair: final descriptors and both scores"""
    cfg = blocks.indoor_config(first_feats_dim=32, gnn_feats_dim=64, dgcnn_k=10, num_head=4, nets=["self", "cross", "self"])
    net = architectures.KPFCNN(cfg).to(DEV)
    sd = {k[6:]: torch.from_numpy(gnn[k]) for k in gnn.files if k.startswith("netsd_")}
    missing, unexpected = net.load_state_dict(sd, strict=True)
    assert not missing and not unexpected                   # same parameter names and shapes as the reference
    batch = _geom(blk)
    batch["features"] = torch.ones(batch["points"][0].shape[0], 1, device=DEV)
    res = net(batch)
    assert res["feats_f"].shape == gnn["net_feats_f"].shape
    assert _err(res["feats_f"], gnn["net_feats_f"]) < TOL
    assert _err(res["scores_overlap"], gnn["net_scores_overlap"]) < TOL
    assert _err(res["scores_saliency"], gnn["net_scores_saliency"]) < TOL


def test_whole_network_stacked_pairs_equal_single_pairs():
    """descriptors of a pair do not depend on what else is stacked in the batch (per-pair statistics, per-pair attention)"""
    cfg = blocks.indoor_config(first_feats_dim=32, gnn_feats_dim=64)
    limits = [30, 28, 28, 30]
    torch.manual_seed(0)
    net = architectures.KPFCNN(cfg).to(DEV)
    for m in net.modules():
        if isinstance(m, blocks.KPConv):
            m.set_kernel_points(torch.randn(15, 3) * 0.4 * m.radius)
    pairs = [synthetic.match3d_pair(s, n_target=1500 + 200 * s)[:2] for s in range(2)]
    singles = []
    for src, tgt in pairs:
        b = dataloader.build_pyramid(np.concatenate([src, tgt]), np.array([len(src), len(tgt)], np.int32), cfg, limits, device=DEV)
        b["features"] = torch.ones(len(src) + len(tgt), 1, device=DEV)
        singles.append({k: v.cpu().numpy() for k, v in net(b).items()})
    pts = np.concatenate([np.concatenate(p) for p in pairs])
    lens = np.array([len(c) for p in pairs for c in p], np.int32)
    b = dataloader.build_pyramid(pts, lens, cfg, limits, device=DEV)
    b["features"] = torch.ones(len(pts), 1, device=DEV)
    res = {k: v.cpu().numpy() for k, v in net(b).items()}
    seg = b["pair_segments"][0].cpu().numpy()
    for k, s in enumerate(singles):
        for name in ("feats_f", "scores_overlap", "scores_saliency"):
            part = res[name][seg[k]:seg[k + 1]]
            assert part.shape == s[name].shape
            assert np.abs(part - s[name]).max() <= 5e-4 * max(np.abs(s[name]).max(), 1e-6), name


def test_point2node_and_node_visibility_vs_reference_golden():
    """'next' row 1 (second half): datasets/dataloader.py:91-198 -- assignment exact, visibility ratios exact"""
    g = np.load(os.path.join(G, "point2node_ref.npz"))
    sv, tv, si, ti = dataloader.point2node_correspondences(_d(g["src_nodes"]), _d(g["src_points"]), _d(g["tgt_nodes"]), _d(g["tgt_points"]),
                                                           torch.from_numpy(g["corr"]))
    assert si.dtype == torch.int64 and np.array_equal(si.cpu().numpy(), g["src_idx"]) and np.array_equal(ti.cpu().numpy(), g["tgt_idx"])
    assert np.array_equal(sv.cpu().numpy(), g["src_node_vis"]) and np.array_equal(tv.cpu().numpy(), g["tgt_node_vis"])
    assert np.array_equal(dataloader.point2node(_d(g["src_nodes"]), _d(g["src_points"])).cpu().numpy(), g["src_idx"])


def test_edge_cases_small_clouds_and_empty_inputs():
    """a cloud with fewer than k+1 nodes repeats the query in its kNN list (the reference's topk would raise); empty inputs
    are no-ops"""
    pts = torch.tensor([[0., 0, 0], [1, 0, 0], [0, 2, 0], [5, 5, 5], [5, 5, 6]], device=DEV)
    idx = ops.knn(pts, ops.cloud_starts([3, 2]), 3).cpu().numpy()
    assert idx[0].tolist() == [1, 2, 0] and idx[3].tolist() == [4, 3, 3] and idx[4].tolist() == [3, 4, 4]
    empty = torch.zeros((0, 3), device=DEV)
    assert ops.knn(empty, ops.cloud_starts([0]), 4).shape == (0, 4)
    assert ops.l2_normalize(torch.zeros((0, 8), device=DEV)).shape == (0, 8)
    assert ops.bias_act(torch.zeros((0, 8), device=DEV), torch.zeros(8, device=DEV), 0.0).shape == (0, 8)
    z = ops.l2_normalize(torch.zeros((3, 8), device=DEV))              # F.normalize eps: zero rows stay zero
    assert float(z.abs().max()) == 0.0
