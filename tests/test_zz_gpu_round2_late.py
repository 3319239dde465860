"""GPU tests written AFTER round 2's GPU budget was spent: none of them had run on a B200 when they were committed, so the file
sorts last on purpose (pytest -x reaches it after everything that has been green on the GPU).

* subsample_batch(features=, classes=) -- grid_subsampling.cpp:34-102, wrapper.cpp:103-326 -- through the host C-ABI entry (the
  reference's module mirror) and the device entry, bit for bit against the CPU oracle (oracle/port.c, pinned to the unmodified
  reference core by tests/test_label_vote_host.py; the kernels' thread bodies are run on the CPU by
  tests/test_subsample_extras_host.py)
* collate_fn_descriptor / batch_grid_subsampling_kpconv / batch_neighbors_kpconv against the golden made by executing the reference's
  own source (tests/golden/make_golden_callsites.py)
* regression test of the fused KPConv's tile ring under ragged neighbourhoods"""
import numpy as np
import pytest
import torch

import oracle
import test_gpu_kpconv_fused as kf
from pcrcg_b200 import dataloader, ops
from pcrcg_b200.cpp_wrappers.cpp_subsampling import grid_subsampling as cpp_subsampling

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _case(seed, nb, n_per, extent, fdim, ldim, n_labels):
    rng = np.random.default_rng(seed)
    lens = rng.integers(max(1, n_per // 2), n_per + 1, size=nb).astype(np.int32)
    n = int(lens.sum())
    pts = (rng.random((n, 3)) * extent).astype(np.float32)
    f = rng.standard_normal((n, fdim)).astype(np.float32)
    c = rng.integers(-2, n_labels - 2, size=(n, ldim)).astype(np.int32)
    return pts, lens, f, c


def _same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        y = y.cpu().numpy() if torch.is_tensor(y) else y
        assert x.dtype == y.dtype and x.shape == y.shape and np.array_equal(x, y)


@pytest.mark.parametrize("seed,nb,n_per,extent,dl,n_labels", [
    (0, 3, 4000, 1.0, 0.1, 4),          # a few points and labels per voxel: frequent two-way ties
    (1, 2, 3000, 0.3, 0.1, 40),         # crowded voxels: up to ~40 distinct labels, the 13 -> 29 -> 59 bucket rehashes
    (2, 1, 500, 1.0, 0.05, 3),          # one cloud, mostly single-point voxels
    (3, 4, 2000, 2.0, 0.25, 12),
])
def test_features_and_classes_match_the_reference(port, seed, nb, n_per, extent, dl, n_labels):
    pts, lens, f, c = _case(seed, nb, n_per, extent, 5, 1, n_labels)
    for kw in (dict(features=f), dict(classes=c), dict(features=f, classes=c), dict(classes=c[:, 0])):
        want = port.subsample_batch_ex(pts, lens, sampleDl=dl, **kw)
        _same(want, cpp_subsampling.subsample_batch(pts, lens, sampleDl=dl, **kw))                      # host buffers, C ABI
        dkw = {k: _t(v) for k, v in kw.items()}
        _same(want, ops.subsample_batch_ex(_t(pts), _t(lens), dl, **dkw))                               # device entry
    want = port.subsample_batch_ex(pts, lens, features=f, classes=c, sampleDl=dl)
    _same(want, dataloader.batch_grid_subsampling_kpconv(_t(pts), _t(lens), features=_t(f), labels=_t(c), sampleDl=dl))


def test_points_and_lens_unchanged_by_the_extras(port):
    pts, lens, f, c = _case(7, 3, 3000, 1.0, 3, 1, 5)
    gp, gl = cpp_subsampling.subsample_batch(pts, lens, sampleDl=0.1)
    ep, el, _, _ = cpp_subsampling.subsample_batch(pts, lens, features=f, classes=c, sampleDl=0.1)
    assert np.array_equal(gp, ep) and np.array_equal(gl, el)


def test_max_p_truncates_features_and_classes_alike(port):
    pts, lens, f, c = _case(4, 3, 3000, 1.0, 2, 1, 6)
    want = port.subsample_batch_ex(pts, lens, features=f, classes=c, sampleDl=0.1, max_p=57)
    assert want[1].tolist() == [57, 57, 57]
    _same(want, cpp_subsampling.subsample_batch(pts, lens, features=f, classes=c, sampleDl=0.1, max_p=57))


def test_multi_column_classes_single_cloud(port):
    pts, lens, f, c = _case(5, 1, 5000, 1.0, 1, 3, 7)
    want = port.subsample_batch_ex(pts, lens, classes=c, sampleDl=0.1)
    _same(want, cpp_subsampling.subsample_batch(pts, lens, classes=c, sampleDl=0.1))
    sp, sc = cpp_subsampling.subsample(pts, classes=c, sampleDl=0.1)                                   # wrapper.cpp:546-553
    assert np.array_equal(sp, want[0]) and np.array_equal(sc, want[2])


def test_all_points_in_one_voxel_many_tied_labels(port):
    """every label occurs once: the answer is purely the container's iteration order, through three rehashes"""
    rng = np.random.default_rng(6)
    for D in (1, 2, 13, 14, 29, 30, 59, 60, 64):
        pts = (rng.random((D, 3)) * 0.01).astype(np.float32)
        c = (rng.permutation(1000)[:D] - 500).astype(np.int32)
        want = port.subsample_batch_ex(pts, [D], classes=c, sampleDl=1.0)
        assert len(want[0]) == 1
        _same(want, cpp_subsampling.subsample_batch(pts, [D], classes=c, sampleDl=1.0))


def test_voxel_counts_around_the_rehash_schedule(port):
    """clouds with 1 ... 300 voxels: k_order leaves the final list in seqA or seqB depending on the number of rehash epochs
    (13 / 29 / 59 / 127 / 257 buckets); the gather of features and classes must follow it"""
    rng = np.random.default_rng(99)
    for M in (1, 2, 12, 13, 14, 28, 29, 30, 59, 60, 127, 128, 257, 258, 300):
        g = int(np.ceil(M ** (1 / 3))) + 1
        cells = rng.permutation(g ** 3)[:M]
        centres = np.stack([cells % g, (cells // g) % g, cells // (g * g)], 1).astype(np.float32) + 0.5
        pts = np.repeat(centres, 3, axis=0) + rng.uniform(-0.3, 0.3, size=(3 * M, 3)).astype(np.float32)
        pts = pts[rng.permutation(len(pts))].astype(np.float32)
        lens = np.array([len(pts)], np.int32)
        f = rng.standard_normal((len(pts), 2)).astype(np.float32)
        c = rng.integers(0, 3, size=(len(pts), 1)).astype(np.int32)
        want = port.subsample_batch_ex(pts, lens, features=f, classes=c, sampleDl=1.0)
        assert len(want[0]) == M
        _same(want, cpp_subsampling.subsample_batch(pts, lens, features=f, classes=c, sampleDl=1.0))
    # and two clouds of different parity in one batch
    a = (rng.random((40, 3)) * 3).astype(np.float32)
    b = (rng.random((400, 3)) * 6).astype(np.float32)
    pts, lens = np.concatenate([a, b]), np.array([40, 400], np.int32)
    f = rng.standard_normal((440, 3)).astype(np.float32)
    c = rng.integers(0, 4, size=(440, 1)).astype(np.int32)
    _same(port.subsample_batch_ex(pts, lens, features=f, classes=c, sampleDl=1.0),
          cpp_subsampling.subsample_batch(pts, lens, features=f, classes=c, sampleDl=1.0))


def test_more_than_64_distinct_labels_in_a_voxel_is_reported():
    pts = np.zeros((65, 3), np.float32)
    with pytest.raises(RuntimeError, match="more than 64 distinct labels"):
        cpp_subsampling.subsample_batch(pts, [65], classes=np.arange(65, dtype=np.int32), sampleDl=1.0)
    with pytest.raises(RuntimeError, match="more than 64 distinct labels"):
        ops.subsample_batch_ex(_t(pts), _t(np.array([65], np.int32)), 1.0, classes=_t(np.arange(65, dtype=np.int32)))


def test_collate_returns_the_reference_dict_for_one_pair():
    """collate_fn_descriptor (datasets/dataloader.py:203-400) for a single pair: the pyramid keys plus everything else the
    reference's dict holds -- pass-through keys untouched, node labels of the coarsest level equal to the direct call"""
    from pcrcg_b200 import blocks, pipeline, synthetic
    src, tgt, _ = synthetic.match3d_pair(3, n_target=3000)
    rng = np.random.default_rng(3)
    corr = torch.from_numpy(np.stack([rng.integers(0, len(src), 700), rng.integers(0, len(tgt), 700)], 1))
    vm = object()
    item = dict(src_pcd=src, tgt_pcd=tgt, src_feats=np.ones((len(src), 1), np.float32), tgt_feats=np.ones((len(tgt), 1), np.float32),
                rot=np.eye(3), trans=np.zeros((3, 1)), correspondences=corr, sample="7-scenes-redkitchen@0", src_valid_map1=vm,
                not_a_reference_key=1)
    cfg, limits = blocks.indoor_config(), pipeline.CALIBRATED_LIMITS["3dmatch_synthetic"]
    b = dataloader.collate_fn_descriptor([item], cfg, limits, device=DEV)
    want = {"points", "neighbors", "pools", "upsamples", "features", "stack_lengths", "rot", "trans", "correspondences", "src_pcd_raw",
            "tgt_pcd_raw", "sample", "node_overlap_gt", "points2node", "src_valid_map1"}
    assert want <= set(b) and "not_a_reference_key" not in b
    assert b["correspondences"] is corr and b["src_valid_map1"] is vm and b["sample"] == "7-scenes-redkitchen@0"
    assert torch.equal(b["rot"], torch.eye(3, dtype=torch.float64)) and tuple(b["trans"].shape) == (3, 1)
    assert np.array_equal(b["src_pcd_raw"].numpy(), src.astype(np.float32)) and b["tgt_pcd_raw"].dtype == torch.float32
    assert len(b["points"]) == 4 and b["features"].shape == (len(src) + len(tgt), 1)
    nodes, n_src = b["points"][-1], int(b["stack_lengths"][-1][0])
    sv, tv, s2n, t2n = dataloader.point2node_correspondences(nodes[:n_src], src, nodes[n_src:], tgt, corr)
    assert torch.equal(b["node_overlap_gt"], torch.cat((sv, tv))) and torch.equal(b["points2node"], torch.cat((s2n, t2n)))
    assert b["node_overlap_gt"].shape[0] == nodes.shape[0] and b["points2node"].shape[0] == len(src) + len(tgt)
    assert float(b["node_overlap_gt"].min()) >= 0.0 and float(b["node_overlap_gt"].max()) <= 1.0
    # several pairs stacked (not a reference case): only the pyramid keys
    b2 = dataloader.collate_fn_descriptor([item, item], cfg, limits, device=DEV)
    assert "rot" not in b2 and b2["stack_lengths"][0].numel() == 4


def test_dataloader_mirrors_accept_the_references_cpu_inputs(port):
    """datasets/dataloader.py:273-301 hands batch_grid_subsampling_kpconv / batch_neighbors_kpconv CPU tensors; the mirrors copy
    them to the GPU and return device tensors with the reference's values"""
    pts, lens, _, _ = _case(11, 2, 3000, 1.0, 1, 1, 3)
    sp, sl = dataloader.batch_grid_subsampling_kpconv(torch.from_numpy(pts), torch.from_numpy(lens), sampleDl=0.1)
    op, ol = port.subsample_batch(pts, lens, 0.1)
    assert sp.is_cuda and sl.is_cuda and np.array_equal(sp.cpu().numpy(), op) and np.array_equal(sl.cpu().numpy(), ol)
    rows = dataloader.batch_neighbors_kpconv(op, pts, ol, lens.tolist(), 0.2, 30)           # NumPy arrays and a list
    want = port.batch_query(op, pts, ol, lens, 0.2, limit=30)
    assert rows.is_cuda and rows.dtype == torch.int32 and np.array_equal(rows.cpu().numpy(), want)


# ---- the reference's own call sites (datasets/dataloader.py), golden = their SOURCE executed over the reference core ------------
import os                                                                                      # noqa: E402

_G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "callsites_ref.npz"))


def test_subsampling_call_site_vs_reference_source():
    """batch_grid_subsampling_kpconv (datasets/dataloader.py:14-52): every branch, CPU tensors in as the collate passes them"""
    P, B = torch.from_numpy(_G["sub_points"]), torch.from_numpy(_G["sub_lens"])
    f, c = torch.from_numpy(_G["sub_features"]), torch.from_numpy(_G["sub_labels"])
    for tag, kw in (("plain", {}), ("feat", dict(features=f)), ("lab", dict(labels=c)), ("both", dict(features=f, labels=c))):
        out = dataloader.batch_grid_subsampling_kpconv(P, B, sampleDl=0.06, **kw)
        assert len(out) == 2 + len(kw)
        for i, o in enumerate(out):
            assert np.array_equal(o.cpu().numpy(), _G[f"sub_{tag}_{i}"]), (tag, i)
    out = dataloader.batch_grid_subsampling_kpconv(P, B, sampleDl=0.06, max_p=300)
    assert np.array_equal(out[0].cpu().numpy(), _G["sub_maxp_0"]) and np.array_equal(out[1].cpu().numpy(), _G["sub_maxp_1"])


def test_neighbour_call_site_vs_reference_source():
    """batch_neighbors_kpconv (datasets/dataloader.py:54-69): truncated to max_neighbors, and full width for max_neighbors = 0"""
    P, B = torch.from_numpy(_G["sub_points"]), torch.from_numpy(_G["sub_lens"])
    sp, sl = torch.from_numpy(_G["sub_plain_0"]), torch.from_numpy(_G["sub_plain_1"])
    rows = dataloader.batch_neighbors_kpconv(sp, P, sl, B, 0.15, 20)
    assert np.array_equal(rows.cpu().numpy(), _G["nb_pool_20"])
    rows = dataloader.batch_neighbors_kpconv(sp, sp, sl, sl, 0.15, 0)
    assert np.array_equal(rows.cpu().numpy(), _G["nb_conv_full"])


def test_collate_vs_reference_source():
    """collate_fn_descriptor (datasets/dataloader.py:203-400) on one pair: every list of the pyramid, the features and the node
    labels of the coarsest level against the reference's own function"""
    from pcrcg_b200 import blocks
    src, tgt, corr = _G["col_src"], _G["col_tgt"], torch.from_numpy(_G["col_corr"])
    item = dict(src_pcd=src, tgt_pcd=tgt, src_feats=np.ones((len(src), 1), np.float32), tgt_feats=np.ones((len(tgt), 1), np.float32),
                rot=np.eye(3, dtype=np.float32), trans=np.zeros((3, 1), np.float32), correspondences=corr, sample="synthetic@17")
    b = dataloader.collate_fn_descriptor([item], blocks.indoor_config(), _G["col_limits"].tolist(), device=DEV)
    assert set(_G["col_keys"].tolist()) <= set(b)
    for l in range(4):
        assert np.array_equal(b["points"][l].cpu().numpy(), _G[f"col_points_{l}"])
        assert np.array_equal(b["stack_lengths"][l].cpu().numpy(), _G[f"col_stack_lengths_{l}"])
        for k in ("neighbors", "pools", "upsamples"):
            assert np.array_equal(b[k][l].cpu().numpy().astype(np.int64), _G[f"col_{k}_{l}"]), (k, l)
    assert np.array_equal(b["features"].cpu().numpy(), _G["col_features"])
    assert np.array_equal(b["points2node"].cpu().numpy(), _G["col_points2node"])
    assert np.allclose(b["node_overlap_gt"].cpu().numpy(), _G["col_node_overlap_gt"], rtol=0, atol=1e-6)


# ---- fused KPConv: tile ring ----------------------------------------------------------------------------------------------------
def test_fused_tile_ring_under_ragged_neighbourhoods():
    """Regression test of the tile-ring race (tests/test_fused_protocol.py): neighbourhoods cut to random lengths mix
    1-k-step and 4-k-step points, so the 13 producer warps of a CTA drift apart over ~100 tiles each; the kernel must give the
    two-kernel path's result, and the same bits on every repetition."""
    q, s, rows, seg = kf._geometry(21, 20000, 64, n_pairs=3)
    ns, width = s.shape[0], rows.shape[1]
    g = torch.Generator().manual_seed(21)
    keep = torch.randint(1, width + 1, (rows.shape[0], 1), generator=g).to(DEV)
    keep[::5] = width                                                     # every fifth point keeps its full list
    cols = torch.arange(width, device=DEV).view(1, -1)
    rows = torch.where(cols < keep, rows, torch.full_like(rows, ns)).contiguous()
    x = torch.randn(ns, 64, generator=g)
    w = torch.randn(15, 64, 64, generator=g) / np.sqrt(15 * 64)
    kp = torch.randn(15, 3, generator=g) * 0.03
    xp = kf._planes(x.to(DEV))
    from pcrcg_b200._lib import lib, check
    try:
        _ring_body(q, s, rows, xp, kp, w, seg)
    finally:
        check(lib().pcrcg_set_option(b"kpconv_fused", 1))


def _ring_body(q, s, rows, xp, kp, w, seg):
    two, _, _ = kf._run(0, q, s, rows, xp, kp.to(DEV), w.to(DEV), seg)
    first, _, _ = kf._run(1, q, s, rows, xp, kp.to(DEV), w.to(DEV), seg)
    assert kf._err(first, two) < 1e-4, kf._err(first, two)
    first = first.clone()
    for _ in range(15):
        again, _, _ = kf._run(1, q, s, rows, xp, kp.to(DEV), w.to(DEV), seg)
        assert torch.equal(again, first)
