"""CPU: pins the oracle (oracle/port.c, blocks_port.py, projection_port.py) against the golden
vectors produced by the reference itself (tests/golden/make_golden.py) and, where the reference
tree is available, against the reference run live."""
import hashlib
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import blocks_port as bp
from oracle import projection_port as pp

G = os.path.join(os.path.dirname(__file__), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def pre():
    return np.load(os.path.join(G, "preprocess_ref.npz"))


@pytest.fixture(scope="module")
def blk():
    return np.load(os.path.join(G, "blocks_ref.npz"))


def test_bucket_schedule_matches_libstdcxx(port):
    s = port.bucket_schedule()
    assert s[:16].tolist() == [13, 29, 59, 127, 257, 541, 1109, 2357, 5087, 10273, 20753, 42043, 85229, 172933, 351061, 712697]
    if oracle.have_ref():
        r = oracle.ref().bucket_schedule(1_500_000)
        assert r[0] == 1 and r[1:].tolist() == s[:len(r) - 1].tolist()


def test_subsample_small_golden(port, pre):
    pts, lens = pre["small_pts"], pre["small_lens"]
    for dl in (0.1, 0.25):
        sp, sl = port.subsample_batch(pts, lens, dl)
        assert np.array_equal(sl, pre[f"small_sub_{dl}_lens"]) and np.array_equal(sp, pre[f"small_sub_{dl}_pts"])
    sp, sl = port.subsample_batch(pts, lens, 0.1, max_p=100)
    assert np.array_equal(sl, pre["small_sub_maxp_lens"]) and np.array_equal(sp, pre["small_sub_maxp_pts"])


def test_neighbors_small_golden(port, pre):
    pts, lens = pre["small_pts"], pre["small_lens"]
    rows = port.batch_query(pts, pts, lens, lens, 0.2)
    assert np.array_equal(rows, pre["small_nb_canonical"])
    # the raw reference rows hold the same SETS; canonicalising them reproduces the canonical rows
    can, changed = port.canonicalise_rows(pts, pts, pre["small_nb_raw"])
    assert np.array_equal(can, pre["small_nb_canonical"]) and changed == int(pre["small_nb_rows_reordered"])
    assert np.array_equal(np.sort(pre["small_nb_raw"], 1), np.sort(rows, 1))


def test_demo_pair_pyramid_digests(port, pre, demo_pair):
    """Full reference pyramid of the demo pair (limits 38/36/36/38) reproduced by the C port."""
    pts = np.concatenate(demo_pair)
    lens = np.array([len(demo_pair[0]), len(demo_pair[1])], np.int32)
    limits = pre["demo_limits"].tolist()
    r = 0.0625
    sizes = []
    for l in range(4):
        sizes.append(len(pts))
        assert sha(pts) == str(pre[f"demo_points_sha_{l}"])
        if l >= 1:       # level 0 brute force (8.6e8 pairs x 3 calls) is covered on the GPU box; keep the CPU suite short
            assert sha(port.batch_query(pts, pts, lens, lens, r, limits[l])) == str(pre[f"demo_neighbors_sha_{l}"])
        if l < 3:
            p2, l2 = port.subsample_batch(pts, lens, 2 * r / 2.5)
            if l >= 1:
                assert sha(port.batch_query(p2, pts, l2, lens, r, limits[l])) == str(pre[f"demo_pools_sha_{l}"])
                assert sha(port.batch_query(pts, p2, lens, l2, 2 * r, limits[l])) == str(pre[f"demo_upsamples_sha_{l}"])
            pts, lens = p2, l2
        r *= 2
    assert sizes == pre["demo_level_sizes"].tolist() == [39939, 9932, 2612, 758]


def test_order_closed_form_equals_container_simulation(port):
    """SURVEY App. A.2 closed form (what the CUDA kernel implements) == literal linked-list simulation."""
    sched = port.bucket_schedule().astype(np.int64)
    rng = np.random.default_rng(0)
    for n in (1, 13, 14, 30, 500, 6000):
        pts = (rng.random((n * 3, 3)) * 2).astype(np.float32)
        keys, _, _ = port.voxel_keys(pts, 0.11)
        _, first = np.unique(keys, return_index=True)
        U = keys[np.sort(first)]                       # unique keys, first-occurrence order
        M = len(U)
        L = np.zeros(0, np.int64)
        done = 0
        for P in sched:
            hi = min(M, int(P))
            S = np.concatenate([L, np.arange(done, hi)])
            b = (U[S] % np.uint64(P)).astype(np.int64)
            firstpos = np.full(int(P), len(S), np.int64)
            np.minimum.at(firstpos, b, np.arange(len(S)))
            order = np.lexsort((-np.arange(len(S)), -firstpos[b]))
            L = S[order]
            done = hi
            if done >= M:
                break
        sp, _ = port.subsample_batch(pts, [len(pts)], 0.11)
        # recover the container order from the barycentres: recompute the keys of the output points' voxels
        sums = np.zeros((M, 3), np.float32); cnt = np.zeros(M, np.int64)
        rank = {int(k): i for i, k in enumerate(U)}
        for p, k in zip(pts, keys):
            i = rank[int(k)]
            sums[i] += p; cnt[i] += 1
        bary = sums * (1.0 / cnt).astype(np.float32)[:, None]
        assert np.array_equal(sp, bary[L])


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
def test_port_equals_reference_live(port):
    R = oracle.ref()
    rng = np.random.default_rng(3)
    for trial in range(4):
        clouds = [rng.normal(size=(int(rng.integers(1, 3000)), 3)) * rng.uniform(0.1, 3) for _ in range(3)]
        if trial % 2:
            clouds = [np.round(c / 0.01) * 0.01 for c in clouds]
        pts = np.concatenate(clouds).astype(np.float32)
        lens = np.array([len(c) for c in clouds], np.int32)
        for dl in (0.05, 0.3):
            a, al = R.subsample_batch(pts, lens, dl)
            b, bl = port.subsample_batch(pts, lens, dl)
            assert np.array_equal(al, bl) and np.array_equal(a, b)
        rows, _ = oracle.ref_batch_query_canonical(pts, pts, lens, lens, 0.25)
        assert np.array_equal(rows, port.batch_query(pts, pts, lens, lens, 0.25))


# ---------------------------------------------------------------------------------------------
def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _close(a, b, tol=2e-5):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-30)


def test_blocks_port_vs_reference_golden(blk):
    P = [_t(blk[f"points_{l}"]) for l in range(4)]
    nb = [_t(blk[f"neighbors_{l}"]) for l in range(4)]
    pools = [_t(blk[f"pools_{l}"]) for l in range(4)]
    ups = [_t(blk[f"upsamples_{l}"]) for l in range(4)]
    out = bp.kpconv(P[0], P[0], nb[0], _t(blk["kpconv_x"]), _t(blk["kpconv_kp"]), _t(blk["kpconv_w"]), 0.05)
    assert _close(out, blk["kpconv_out"])
    x1 = torch.ones(P[0].shape[0], 1)
    out = bp.kpconv(P[1], P[0], pools[0], x1, _t(blk["kpconv1_kp"]), _t(blk["kpconv1_w"]), 0.05)
    assert _close(out, blk["kpconv1_out"])
    xs = bp.simple_block(x1, P[0], P[0], nb[0], dict(kernel_points=_t(blk["simple_kp"]), weights=_t(blk["simple_w"]), KP_extent=0.05))
    assert _close(xs, blk["simple_out"])
    p = dict(unary1=_t(blk["rb_unary1.mlp.weight"]), kernel_points=_t(blk["rb_KPConv.kernel_points"]),
             weights=_t(blk["rb_KPConv.weights"]), KP_extent=0.05, unary2=_t(blk["rb_unary2.mlp.weight"]),
             shortcut=_t(blk["rb_unary_shortcut.mlp.weight"]))
    xr = bp.resnetb_block(_t(blk["simple_out"]), P[0], P[0], nb[0], p, strided=False)
    assert _close(xr, blk["rb_out"])
    p = dict(unary1=_t(blk["rs_unary1.mlp.weight"]), kernel_points=_t(blk["rs_KPConv.kernel_points"]),
             weights=_t(blk["rs_KPConv.weights"]), KP_extent=0.05, unary2=_t(blk["rs_unary2.mlp.weight"]), shortcut=None)
    xo = bp.resnetb_block(_t(blk["rb_out"]), P[1], P[0], pools[0], p, strided=True)
    assert _close(xo, blk["rs_out"])
    assert np.array_equal(bp.max_pool(_t(blk["pool_x"]), pools[0]).numpy(), blk["max_pool_out"])
    assert np.array_equal(bp.closest_pool(_t(blk["closest_x"]), ups[0]).numpy(), blk["closest_pool_out"])


def test_encoder_port_vs_reference_golden(blk):
    enc = np.load(os.path.join(G, "encoder_ref.npz"))
    sd = {k[3:]: enc[k] for k in enc.files if k.startswith("sd_")}
    blocks = bp.encoder_blocks_from_state_dict(sd)
    batch = {k: [_t(blk[f"{k}_{l}"]) for l in range(4)] for k in ("points", "neighbors", "pools", "upsamples")}
    x, outs = bp.encoder(torch.ones(batch["points"][0].shape[0], 1), batch, blocks)
    for i, o in enumerate(outs):
        assert abs(float(o.abs().max()) - float(enc[f"block_out_absmax_{i}"])) <= 1e-3 * float(enc[f"block_out_absmax_{i}"])
    assert _close(x, enc["encoder_out"], 1e-4)


def test_projection_port_vs_reference_golden():
    g = np.load(os.path.join(G, "projection_ref.npz"))
    views = []
    for vi in (0, 1):
        i2, i3 = pp.projection(g["points"], g[f"v{vi}_depth"], g[f"v{vi}_world2camera"], g[f"v{vi}_intrinsics"])
        assert np.array_equal(i2, g[f"v{vi}_inds2d"]) and np.array_equal(i3, g[f"v{vi}_inds3d"])
        views.append((g[f"v{vi}_feature2d"], g[f"v{vi}_valid_map"], i2, i3))
    x = pp.scatter_image_features(len(g["points"]), [views[1], views[0]])
    assert np.array_equal(x, g["x_out"])


def test_decoder_port_vs_reference_golden(blk):
    dec = np.load(os.path.join(G, "decoder_ref.npz"))
    batch = {k: [_t(blk[f"{k}_{l}"]) for l in range(4)] for k in ("points", "upsamples")}
    skips = [_t(dec[f"skip_{i}"]) for i in range(3)]
    Ws = [_t(dec["sd_1.mlp.weight"]), _t(dec["sd_3.mlp.weight"]), _t(dec["sd_5.mlp.weight"])]
    feats, so, ss, raw = bp.decoder(_t(dec["bottleneck_x"]), skips[:3], batch, Ws, 32)
    assert _close(raw, dec["decoder_out"], 1e-4) and _close(feats, dec["feats_f"], 1e-4)
    assert _close(so, dec["scores_overlap"], 1e-4) and _close(ss, dec["scores_saliency"], 1e-4)


# ---- bottleneck GNN + whole network ------------------------------------
from oracle import gcn_port as gp  # noqa: E402


@pytest.fixture(scope="module")
def gnn():
    return np.load(os.path.join(G, "gnn_ref.npz"))


def test_gcn_port_vs_reference_golden(gnn):
    sd = {k[6:]: _t(gnn[k]) for k in gnn.files if k.startswith("gcnsd_")}
    coords, lens, feats = _t(gnn["gcn_coords"]), gnn["gcn_lens"], _t(gnn["gcn_feats"])
    assert np.array_equal(gp.knn_lists(coords, lens, 10).numpy(), gnn["gcn_knn"])
    out = gp.gcn(coords, lens, feats, sd, ["self", "cross", "self"], 4, 10)
    assert _close(out, gnn["gcn_out"], 1e-4)


def test_whole_network_port_vs_reference_golden(blk, gnn):
    """encoder -> bottle -> GNN -> saliency -> decoder, every stage from the oracle ports, vs KPFCNN.forward of the reference"""
    sd = {k[6:]: _t(gnn[k]) for k in gnn.files if k.startswith("netsd_")}
    blocks = bp.encoder_blocks_from_state_dict(sd, prefix="encoder_blocks.")
    batch = {k: [_t(blk[f"{k}_{l}"]) for l in range(4)] for k in ("points", "neighbors", "pools", "upsamples")}
    x, outs = bp.encoder(torch.ones(batch["points"][0].shape[0], 1), batch, blocks)
    xb = gp.bottleneck(x, batch["points"][3], blk["stack_lengths_3"], sd, ["self", "cross", "self"], 4, 10)
    Ws = [sd["decoder_blocks.1.mlp.weight"], sd["decoder_blocks.3.mlp.weight"], sd["decoder_blocks.5.mlp.weight"]]
    feats, so, ss, _ = bp.decoder(xb, [outs[1], outs[4], outs[7]], batch, Ws, 32)
    assert _close(feats, gnn["net_feats_f"], 1e-3)
    assert _close(so, gnn["net_scores_overlap"], 1e-3) and _close(ss, gnn["net_scores_saliency"], 1e-3)


def test_point2node_port_vs_reference_golden():
    g = np.load(os.path.join(G, "point2node_ref.npz"))
    sv, tv, si, ti = gp.point2node_correspondences(_t(g["src_nodes"]), _t(g["src_points"]), _t(g["tgt_nodes"]), _t(g["tgt_points"]),
                                                   torch.from_numpy(g["corr"]))
    assert np.array_equal(si.numpy(), g["src_idx"]) and np.array_equal(ti.numpy(), g["tgt_idx"])
    assert np.array_equal(sv.numpy(), g["src_node_vis"]) and np.array_equal(tv.numpy(), g["tgt_node_vis"])


def test_matching_port_vs_reference_golden():
    from oracle import matching_port as mp
    g = np.load(os.path.join(G, "matching_ref.npz"))
    r, c = mp.mutual_matches(g["src_feat"], g["tgt_feat"])
    assert np.array_equal(r, g["mutual_rows"]) and np.array_equal(c, g["mutual_cols"])
    wo, w = mp.inlier_ratios(g["src_pcd"], g["tgt_pcd"], g["src_feat"], g["tgt_feat"], g["rot"], g["trans"])
    assert abs(wo - float(g["inlier_ratio_wo"])) < 1e-6 and abs(w - float(g["inlier_ratio_w"])) < 1e-6


def test_cpu_pyramid_equals_the_references_own_collate():
    """oracle/checks.cpu_pyramid (the checker behind tests/test_gpu_benchsize.py and bench.py's parity_check) against
    tests/golden/callsites_ref.npz = the output of the reference's collate_fn_descriptor SOURCE executed unchanged over the
    unmodified reference core (tests/golden/make_golden_callsites.py)"""
    from oracle import checks
    g = np.load(os.path.join(G, "callsites_ref.npz"))
    pyr = checks.cpu_pyramid(g["col_src"], g["col_tgt"], g["col_limits"].tolist(), 0.025, 2.5, 4)
    for l in range(4):
        assert np.array_equal(pyr["points"][l], g[f"col_points_{l}"])
        assert np.array_equal(pyr["stack_lengths"][l], g[f"col_stack_lengths_{l}"])
        for k in ("neighbors", "pools", "upsamples"):
            assert np.array_equal(pyr[k][l].astype(np.int64), g[f"col_{k}_{l}"]), (k, l)
    assert int(g["rows_canonicalised"]) > 0          # the golden does contain tied rows


def test_port_equals_the_references_subsampling_call_site(port):
    """datasets/dataloader.py:14-52 executed from the reference's source (all four branches + max_p) vs the C port"""
    g = np.load(os.path.join(G, "callsites_ref.npz"))
    pts, lens, f, c = g["sub_points"], g["sub_lens"], g["sub_features"], g["sub_labels"]
    for tag, kw in (("plain", {}), ("feat", dict(features=f)), ("lab", dict(classes=c)), ("both", dict(features=f, classes=c))):
        out = port.subsample_batch_ex(pts, lens, sampleDl=0.06, **kw)
        for i, o in enumerate(out):
            assert np.array_equal(o, g[f"sub_{tag}_{i}"]), (tag, i)
    out = port.subsample_batch(pts, lens, 0.06, 300)
    assert np.array_equal(out[0], g["sub_maxp_0"]) and np.array_equal(out[1], g["sub_maxp_1"])
