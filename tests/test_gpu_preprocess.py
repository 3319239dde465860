"""GPU parity: grid subsampling and radius search through the C ABI vs the CPU oracle.
Bit-exact bar: identical s_len, identical output ORDER, barycentres max-abs-diff == 0 (<= 1e-6 required),
identical neighbour rows (int32) under the canonical (d2, index) tie rule."""
import numpy as np
import pytest
import torch

import oracle
from pcrcg_b200 import ops, synthetic
from pcrcg_b200.cpp_wrappers.cpp_subsampling import grid_subsampling as cpp_subsampling
from pcrcg_b200.cpp_wrappers.cpp_neighbors import radius_neighbors as cpp_neighbors

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _stack(clouds):
    return np.concatenate(clouds).astype(np.float32), np.array([len(c) for c in clouds], np.int32)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dl", [0.05, 0.1, 0.2, 0.031, 1.7])
def test_subsample_demo_pair(demo_pair, port, dl):
    pts, lens = _stack(demo_pair)
    op, ol = port.subsample_batch(pts, lens, dl)
    gp, gl = cpp_subsampling.subsample_batch(pts, lens, sampleDl=dl)       # host C-ABI entry
    assert gl.dtype == np.int32 and gp.dtype == np.float32
    assert np.array_equal(gl, ol)
    assert np.abs(gp - op).max() <= 1e-6
    assert np.array_equal(gp, op), "order or barycentre bits differ"
    dp, dlens = ops.subsample_batch(_t(pts), _t(lens), dl)                  # device entry
    assert np.array_equal(dp.cpu().numpy(), op) and np.array_equal(dlens.cpu().numpy(), ol)


def test_subsample_anchor_sizes(demo_pair):
    # survey-verified anchors of the reference on cloud_bin_21 alone
    src = demo_pair[0]
    for dl, m in ((0.05, 6178), (0.1, 1573), (0.2, 450)):
        gp, gl = cpp_subsampling.subsample_batch(src, [len(src)], sampleDl=dl)
        assert gl.tolist() == [m] and gp.shape == (m, 3)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_subsample_synthetic_many_clouds(port, seed):
    rng = np.random.default_rng(seed)
    clouds = []
    for k in range(7):
        n = int(rng.integers(1, 4000))
        c = rng.normal(size=(n, 3)) * rng.uniform(0.05, 2.0) + rng.uniform(-5, 5, size=3)
        if k % 3 == 0:
            c = np.round(c / 0.006) * 0.006          # lattice -> many points per voxel, exact ties
        clouds.append(c)
    clouds.append(np.array([[0.1, 0.2, 0.3]]))                    # 1-point cloud
    clouds.append(np.repeat(np.array([[1.0, -2.0, 3.0]]), 50, 0))  # duplicates
    pts, lens = _stack(clouds)
    for dl in (0.04, 0.13):
        for max_p in (0, 17):
            op, ol = port.subsample_batch(pts, lens, dl, max_p)
            gp, gl = cpp_subsampling.subsample_batch(pts, lens, sampleDl=dl, max_p=max_p)
            assert np.array_equal(gl, ol)
            assert np.array_equal(gp, op)


def test_subsample_big_voxels(port):
    # few huge voxels: long sequential sums, tiny M
    rng = np.random.default_rng(5)
    pts = (rng.random((30000, 3)) * 2.0).astype(np.float32)
    op, ol = port.subsample_batch(pts, [30000], 1.0)
    gp, gl = cpp_subsampling.subsample_batch(pts, [30000], sampleDl=1.0)
    assert np.array_equal(gl, ol) and np.array_equal(gp, op)


def test_subsample_kitti_shaped(port):
    a, b, _ = synthetic.kitti_pair(0)
    pts, lens = _stack([a, b])
    op, ol = port.subsample_batch(pts, lens, 0.3)
    gp, gl = cpp_subsampling.subsample_batch(pts, lens, sampleDl=0.3)
    assert np.array_equal(gl, ol) and np.array_equal(gp, op)


def test_subsample_idempotent_at_full_size():
    # size-independent property at bench size: subsampling an already subsampled cloud with the
    # same dl leaves every barycentre inside its own voxel -> same count
    src, tgt, _ = synthetic.match3d_pair(3)
    pts, lens = _stack([src, tgt] * 8)
    p1, l1 = ops.subsample_batch(_t(pts), _t(lens), 0.05)
    p2, l2 = ops.subsample_batch(p1, l1, 0.05)
    assert l1.sum().item() == p1.shape[0]
    assert (l2 <= l1).all()
    # the 8 copies of the same pair give identical results (clouds are processed independently)
    l = l1.cpu().numpy().reshape(8, 2)
    assert (l == l[0]).all()
    o = np.r_[0, np.cumsum(l1.cpu().numpy())]
    p1c = p1.cpu().numpy()
    for k in range(1, 8):
        assert np.array_equal(p1c[o[2 * k]:o[2 * k + 2]], p1c[o[0]:o[2]])


def test_subsample_errors():
    with pytest.raises(RuntimeError, match=r"points.shape is not \(N, 3\)"):
        cpp_subsampling.subsample_batch(np.zeros((5, 2), np.float32), [5])
    with pytest.raises(RuntimeError, match="Error"):
        cpp_subsampling.subsample_batch(np.zeros((0, 3), np.float32), [0])
    with pytest.raises(TypeError):
        cpp_subsampling.subsample_batch(np.zeros((5, 3), np.float32), [5], 0.1)     # sampleDl is keyword-only


# ---------------------------------------------------------------------------------------------
def _pyramid_level(port, pts, lens, dl):
    return port.subsample_batch(pts, lens, dl)


def test_radius_demo_pair_level1_all_three_calls(demo_pair, port):
    pts0, l0 = _stack(demo_pair)
    p1, l1 = port.subsample_batch(pts0, l0, 0.05)
    p2, l2 = port.subsample_batch(p1, l1, 0.1)
    r = 0.125
    for (q, ql, s, sl, rad) in ((p1, l1, p1, l1, r), (p2, l2, p1, l1, r), (p1, l1, p2, l2, 2 * r)):
        o = port.batch_query(q, s, ql, sl, rad)
        g = cpp_neighbors.batch_query(q, s, ql, sl, radius=rad)
        assert g.dtype == np.int32 and g.shape == o.shape
        assert np.array_equal(g, o)
        for limit in (1, 20, 36):
            gt = ops.batch_query(_t(q), _t(s), _t(ql), _t(sl), rad, limit=limit).cpu().numpy()
            assert np.array_equal(gt, o[:, :limit])


def test_radius_vs_real_reference_canonicalised(demo_pair):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    pts0, l0 = _stack(demo_pair)
    p1, l1 = oracle.ref().subsample_batch(pts0, l0, 0.05)
    rows, changed = oracle.ref_batch_query_canonical(p1, p1, l1, l1, 0.125)
    g = cpp_neighbors.batch_query(p1, p1, l1, l1, radius=0.125)
    assert np.array_equal(g, rows)


def test_radius_demo_pair_level0(demo_pair):
    # full-size level 0 against the real reference (kd-tree) where it travelled, else the C port
    pts0, l0 = _stack(demo_pair)
    if oracle.have_ref():
        o, _ = oracle.ref_batch_query_canonical(pts0, pts0, l0, l0, 0.0625)
    else:
        o = oracle.port().batch_query(pts0, pts0, l0, l0, 0.0625)
    g = cpp_neighbors.batch_query(pts0, pts0, l0, l0, radius=0.0625)
    assert g.shape == o.shape and np.array_equal(g, o)
    assert (g[:, 0] == np.arange(len(pts0))).all()        # Q == S: the query itself is first (d2 = 0)


@pytest.mark.parametrize("seed", [0, 1])
def test_radius_adversarial(port, seed):
    rng = np.random.default_rng(seed)
    lat = np.round(rng.random((1500, 3)) * 0.5 / 0.02) * 0.02                # exact-tie lattice + duplicates
    dup = np.repeat(rng.random((40, 3)), 5, 0)
    line = np.stack([np.linspace(0, 1, 300), np.zeros(300), np.zeros(300)], 1)  # points at r boundary
    far = rng.random((200, 3)) * 100.0                                          # sparse, big extent
    one = np.array([[5.0, 5.0, 5.0]])
    s, sl = _stack([lat, dup, line, far, one])
    q, ql = _stack([lat[::3] + 0.01, dup[::2], line[::2] + np.array([0, 0.1, 0]), far[:50] + 0.5, one + 10.0])
    for rad in (0.1, 1.0 / 299.0 * 3, 0.05):
        o = port.batch_query(q, s, ql, sl, rad) if port.radius_counts(q, s, ql, sl, rad)[1] > 0 else None
        if o is None:
            continue
        g = cpp_neighbors.batch_query(q, s, ql, sl, radius=rad)
        assert np.array_equal(g, o)


def test_radius_long_rows_slow_path(port):
    # > 256 neighbours per query: exercises the extraction path and widths > shared capacity
    rng = np.random.default_rng(3)
    s = (rng.random((3000, 3)) * 0.2).astype(np.float32)
    q = s[:200]
    o = port.batch_query(q, s, [200], [3000], 0.08)
    assert o.shape[1] > 256
    g = cpp_neighbors.batch_query(q, s, [200], [3000], radius=0.08)
    assert np.array_equal(g, o)
    gt = ops.batch_query(_t(q), _t(s), _t(np.array([200], np.int32)), _t(np.array([3000], np.int32)), 0.08, limit=40)
    assert np.array_equal(gt.cpu().numpy(), o[:, :40])


def test_radius_counts_and_symmetry_full_size():
    # size-independent properties at bench size: j in N(i) <=> i in N(j) when Q == S; counts == row fill
    src, tgt, _ = synthetic.match3d_pair(1)
    pts, lens = _stack([src, tgt])
    g = ops.RadiusGrid(_t(pts), _t(lens), 0.0625)
    _, counts, mx = g.query(_t(pts), _t(lens), 0)
    w = int(mx.item())
    rows, counts2, _ = g.query(_t(pts), _t(lens), w)
    rows = rows.cpu().numpy(); counts = counts.cpu().numpy()
    n = len(pts)
    assert np.array_equal(counts, counts2.cpu().numpy())
    assert np.array_equal((rows < n).sum(1), counts)
    i = np.repeat(np.arange(n), w)[rows.reshape(-1) < n]
    j = rows.reshape(-1)[rows.reshape(-1) < n]
    a = set(zip(i.tolist(), j.tolist()))
    assert all((y, x) in a for (x, y) in list(a)[:200000])
    # ascending distances inside each row
    d = np.linalg.norm(pts[np.minimum(rows, n - 1)] - pts[:, None, :], axis=2)
    d[rows >= n] = 1e9
    assert (np.diff(d, axis=1) >= -1e-7).all()


@pytest.mark.parametrize("limit", [34, 80, 0])
def test_radius_cell_centric_equals_per_query_kernel(limit):
    """the cell-centric search (default) and the one-warp-per-query kernel give the same rows, counts and max count, for
    Q == S (the support cells are the work units), Q != S (queries binned into the support grid) and both capacity
    configurations (list width <= 48 / wider)"""
    src, tgt, _ = synthetic.match3d_pair(2, n_target=6000)
    pts, lens = _stack([src, tgt, src[:1], tgt[:0] if False else tgt[:3]])
    P, L = _t(pts), _t(lens)
    sub, sl = ops.subsample_batch(P, L, 0.05)
    far = torch.cat([sub[:50] + 3.0, sub[:50] - 3.0])                       # queries outside the supports' bounding box
    cases = [(P, L, P, L, 0.0625), (sub, sl, P, L, 0.0625), (P, L, sub, sl, 0.125),
             (torch.cat([far, sub]), torch.cat([torch.tensor([100], dtype=torch.int32, device=DEV) + sl[:1], sl[1:]]), P, L, 0.0625)]
    for q, ql, s, sl_, rad in cases:
        g = ops.RadiusGrid(s, sl_, rad)
        out = {}
        for mode in (True, False):
            ops.cell_centric(mode)
            try:
                if limit == 0:
                    _, c, mx = g.query(q, ql, 0)
                    rows, c2, _ = g.query(q, ql, int(mx.item()))
                else:
                    rows, c, mx = g.query(q, ql, limit)
                out[mode] = (rows.cpu().numpy(), c.cpu().numpy(), int(mx.item()))
            finally:
                ops.cell_centric(True)
        assert out[True][2] == out[False][2]
        assert np.array_equal(out[True][1], out[False][1])
        assert np.array_equal(out[True][0], out[False][0])


def test_radius_errors():
    z = np.zeros((4, 3), np.float32)
    with pytest.raises(RuntimeError, match=r"query.shape is not \(N, 3\)"):
        cpp_neighbors.batch_query(np.zeros((4, 2), np.float32), z, [4], [4], radius=0.1)
    with pytest.raises(RuntimeError, match="Wrong number of batch elements"):
        cpp_neighbors.batch_query(z, z, [4], [2, 2], radius=0.1)
    with pytest.raises(RuntimeError, match="Error"):
        cpp_neighbors.batch_query(z + 10.0, z, [4], [4], radius=0.1)      # no neighbour anywhere
