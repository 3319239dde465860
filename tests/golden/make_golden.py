"""Generates the golden vectors under tests/golden/ by running the REFERENCE ITSELF in the build
container (needs /root/reference; never runs on the GPU box).  Re-run:  python tests/golden/make_golden.py

  demo_pair_f32.npz      the reference's demo fragments (assets/cloud_bin_{21,34}.pth) as float32
  preprocess_ref.npz     reference C++ core (oracle/_ref) outputs on small clouds + digests on the demo pair
  blocks_ref.npz         reference models/blocks.py outputs (KPConv, Simple/ResnetBottleneck blocks,
                         max_pool, closest_pool) with their parameters, on a small synthetic pyramid
  encoder_ref.npz        reference KPFCNN.encoder_blocks (11 blocks) output on the same pyramid
  projection_ref.npz     reference projection.py + the colour scatter of models/architectures.py
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
OUT = os.path.join(ROOT, "tests", "golden")

import oracle                                     # noqa: E402
from pcrcg_b200 import synthetic                  # noqa: E402


class Cfg(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def indoor_cfg(first_feats_dim=256, in_feats_dim=1):
    from configs.models import architectures
    return Cfg(num_layers=4, in_points_dim=3, first_feats_dim=first_feats_dim, final_feats_dim=32,
               first_subsampling_dl=0.025, in_feats_dim=in_feats_dim, conv_radius=2.5, deform_radius=5.0,
               num_kernel_points=15, KP_extent=2.0, KP_influence="linear", aggregation_mode="sum",
               fixed_kernel_points="center", use_batch_norm=True, batch_norm_momentum=0.02, deformable=False,
               modulated=False, gnn_feats_dim=256, dgcnn_k=10, num_head=4, nets=["self", "cross", "self"],
               image_feature=False, img_num=2, init_mode="pri3d", node_overlap=False, quaternion=False,
               architecture=architectures["indoor"])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build_pyramid(R, pts, lens, limits, dl0=0.025, conv_radius=2.5, canonical=True):
    """The reference collate (datasets/dataloader.py:239-359) driven through the reference core."""
    P = oracle.port()
    r = dl0 * conv_radius
    out = dict(points=[], neighbors=[], pools=[], upsamples=[], stack_lengths=[])
    for layer in range(4):
        def q(qp, sp, ql, sl, rad):
            rows = R.batch_query(qp, sp, ql, sl, rad)
            if canonical:
                rows, _ = P.canonicalise_rows(qp, sp, rows)
            return np.ascontiguousarray(rows[:, :limits[layer]])
        conv = q(pts, pts, lens, lens, r)
        if layer < 3:
            dl = 2 * r / conv_radius
            pp, pl = R.subsample_batch(pts, lens, dl)
            pool = q(pp, pts, pl, lens, r)
            up = q(pts, pp, lens, pl, 2 * r)
        else:
            pp, pl = np.zeros((0, 3), np.float32), np.zeros((0,), np.int32)
            pool = np.zeros((0, 1), np.int32)
            up = np.zeros((0, 1), np.int32)
        out["points"].append(pts); out["neighbors"].append(conv); out["pools"].append(pool)
        out["upsamples"].append(up); out["stack_lengths"].append(lens)
        pts, lens = pp, pl
        r *= 2
    return out


def main():
    os.chdir(REF)
    sys.path.insert(0, REF)
    torch.manual_seed(0)
    np.random.seed(0)
    torch.set_num_threads(4)
    R, P = oracle.ref(), oracle.port()

    # ---- demo pair -----------------------------------------------------------------------------
    a = torch.load(f"{REF}/assets/cloud_bin_21.pth", weights_only=False).astype(np.float32)
    b = torch.load(f"{REF}/assets/cloud_bin_34.pth", weights_only=False).astype(np.float32)
    np.savez_compressed(f"{OUT}/demo_pair_f32.npz", src=a, tgt=b)

    # ---- preprocess --------------------------------------------------------------------------
    g = {}
    rng = np.random.default_rng(7)
    c1 = (rng.normal(size=(700, 3)) * 0.3).astype(np.float32)
    c2 = (np.round(rng.random((500, 3)) / 0.05) * 0.05).astype(np.float32)     # lattice: ties + duplicates
    pts, lens = np.concatenate([c1, c2]), np.array([700, 500], np.int32)
    g["small_pts"], g["small_lens"] = pts, lens
    for dl in (0.1, 0.25):
        sp, sl = R.subsample_batch(pts, lens, dl)
        g[f"small_sub_{dl}_pts"], g[f"small_sub_{dl}_lens"] = sp, sl
    sp, sl = R.subsample_batch(pts, lens, 0.1, max_p=100)
    g["small_sub_maxp_pts"], g["small_sub_maxp_lens"] = sp, sl
    raw = R.batch_query(pts, pts, lens, lens, 0.2)
    can, changed = P.canonicalise_rows(pts, pts, raw)
    g["small_nb_raw"], g["small_nb_canonical"], g["small_nb_rows_reordered"] = raw, can, np.int64(changed)
    # demo pair digests: full reference pyramid, limits of the survey
    dp, dlens = np.concatenate([a, b]), np.array([len(a), len(b)], np.int32)
    limits = [38, 36, 36, 38]
    pyr = build_pyramid(R, dp, dlens, limits)
    g["demo_limits"] = np.array(limits)
    g["demo_level_sizes"] = np.array([len(p) for p in pyr["points"]])
    g["demo_stack_lengths"] = np.stack(pyr["stack_lengths"])
    for l in range(4):
        g[f"demo_points_sha_{l}"] = sha(pyr["points"][l])
        g[f"demo_neighbors_sha_{l}"] = sha(pyr["neighbors"][l])
        g[f"demo_pools_sha_{l}"] = sha(pyr["pools"][l])
        g[f"demo_upsamples_sha_{l}"] = sha(pyr["upsamples"][l])
    for l in range(3):
        raw = R.batch_query(pyr["points"][l], pyr["points"][l], pyr["stack_lengths"][l], pyr["stack_lengths"][l], 0.0625 * 2 ** l)
        g[f"demo_conv_width_{l}"] = np.int64(raw.shape[1])
    np.savez_compressed(f"{OUT}/preprocess_ref.npz", **g)
    print("preprocess:", g["demo_level_sizes"], [int(g[f"demo_conv_width_{l}"]) for l in range(3)])

    # ---- blocks ------------------------------------------------------------------------------
    from models.blocks import KPConv, SimpleBlock, ResnetBottleneckBlock, max_pool, closest_pool
    src, tgt, _ = synthetic.match3d_pair(11, n_target=1400)
    pts, lens = np.concatenate([src, tgt]), np.array([len(src), len(tgt)], np.int32)
    limits = [30, 28, 28, 30]
    pyr = build_pyramid(R, pts, lens, limits)
    batch = {k: [torch.from_numpy(np.ascontiguousarray(x)).long() if k in ("neighbors", "pools", "upsamples")
                 else torch.from_numpy(np.ascontiguousarray(x)) for x in v] for k, v in pyr.items()}
    g = {f"{k}_{l}": pyr[k][l] for k in pyr for l in range(4)}
    g["limits"] = np.array(limits)
    cfg = indoor_cfg(first_feats_dim=64)
    with torch.no_grad():
        n0 = len(pts)
        # KPConv alone, 16 -> 24 channels, mixed-sign features (exercises the neighbour_num rule)
        conv = KPConv(15, 3, 16, 24, 0.05, 0.0625)
        x = torch.randn(n0, 16)
        g["kpconv_x"], g["kpconv_w"], g["kpconv_kp"] = x.numpy(), conv.weights.numpy(), conv.kernel_points.numpy()
        g["kpconv_out"] = conv(batch["points"][0], batch["points"][0], batch["neighbors"][0], x).numpy()
        # strided geometry (pool lists), Cin = 1 with the all-ones features of the real pipeline
        conv1 = KPConv(15, 3, 1, 8, 0.05, 0.0625)
        x1 = torch.ones(n0, 1)
        g["kpconv1_w"], g["kpconv1_kp"] = conv1.weights.numpy(), conv1.kernel_points.numpy()
        g["kpconv1_out"] = conv1(batch["points"][1], batch["points"][0], batch["pools"][0], x1).numpy()
        # SimpleBlock 1 -> 64 (KPConv 1 -> 32)
        sb = SimpleBlock("simple", 1, 64, 0.0625, 0, cfg)
        g["simple_w"], g["simple_kp"] = sb.KPConv.weights.numpy(), sb.KPConv.kernel_points.numpy()
        xs = sb(x1, batch)
        g["simple_out"] = xs.numpy()
        # ResnetBottleneck 32 -> 64 (same resolution, shortcut Linear) and strided 64 -> 64 (max_pool shortcut)
        rb = ResnetBottleneckBlock("resnetb", 32, 64, 0.0625, 0, cfg)
        for k, v in rb.state_dict().items():
            g["rb_" + k] = v.numpy()
        xr = rb(xs, batch)
        g["rb_out"] = xr.numpy()
        rs = ResnetBottleneckBlock("resnetb_strided", 64, 64, 0.0625, 0, cfg)
        for k, v in rs.state_dict().items():
            g["rs_" + k] = v.numpy()
        g["rs_out"] = rs(xr, batch).numpy()
        # pooling helpers
        xf = torch.randn(n0, 20)
        g["pool_x"] = xf.numpy()
        g["max_pool_out"] = max_pool(xf, batch["pools"][0]).numpy()
        xc = torch.randn(len(pyr["points"][1]), 20)
        g["closest_x"] = xc.numpy()
        g["closest_pool_out"] = closest_pool(xc, batch["upsamples"][0]).numpy()
    np.savez_compressed(f"{OUT}/blocks_ref.npz", **g)
    print("blocks: n =", [len(p) for p in pyr["points"]])

    # ---- encoder -----------------------------------------------------------------------------
    from models.architectures import KPFCNN
    cfg = indoor_cfg(first_feats_dim=32)
    torch.manual_seed(1); np.random.seed(1)
    net = KPFCNN(cfg).eval()
    ge = {}
    with torch.no_grad():
        x = torch.ones(len(pts), 1)
        for bi, blk in enumerate(net.encoder_blocks):
            x = blk(x, batch)
            ge[f"block_out_absmax_{bi}"] = np.float32(x.abs().max().item())
        ge["encoder_out"] = x.numpy()
    for k, v in net.encoder_blocks.state_dict().items():
        ge["sd_" + k] = v.numpy()
    np.savez_compressed(f"{OUT}/encoder_ref.npz", **ge)
    print("encoder out", x.shape)

    # ---- decoder tail: models/architectures.py:567-582 ------------------
    import torch.nn.functional as F
    cfg = indoor_cfg(first_feats_dim=32)
    cfg.gnn_feats_dim = 64
    torch.manual_seed(2); np.random.seed(2)
    net = KPFCNN(cfg).eval()
    gd = {}
    with torch.no_grad():
        x = torch.ones(len(pts), 1)
        skip_x = []
        for bi, blk in enumerate(net.encoder_blocks):
            if bi in net.encoder_skips:
                skip_x.append(x)
            x = blk(x, batch)
        for i, t in enumerate(skip_x):
            gd[f"skip_{i}"] = t.numpy()
        xb = torch.randn(x.shape[0], cfg.gnn_feats_dim + 2)          # [scores_c_raw, scores_saliency, feats_gnn_raw]
        gd["bottleneck_x"] = xb.numpy()
        x = xb
        skips = list(skip_x)
        for bi, blk in enumerate(net.decoder_blocks):
            if bi in net.decoder_concats:
                x = torch.cat([x, skips.pop()], dim=1)
            x = blk(x, batch)
        gd["decoder_out"] = x.numpy()
        feats_f = x[:, :cfg.final_feats_dim]
        so = torch.clamp(torch.sigmoid(x[:, cfg.final_feats_dim].view(-1)), min=0, max=1)
        ss = torch.clamp(torch.sigmoid(x[:, cfg.final_feats_dim + 1].view(-1)), min=0, max=1)
        gd["feats_f"] = F.normalize(feats_f, p=2, dim=1).numpy()
        gd["scores_overlap"], gd["scores_saliency"] = net.regular_score(so).numpy(), net.regular_score(ss).numpy()
    for k, v in net.decoder_blocks.state_dict().items():
        gd["sd_" + k] = v.numpy()
    gd["encoder_skip_dims"] = np.array(net.encoder_skip_dims)
    np.savez_compressed(f"{OUT}/decoder_ref.npz", **gd)
    print("decoder out", x.shape)

    # ---- bottleneck GNN: models/gcn.py:37-217, models/architectures.py:528-565 ---------
    from models.gcn import GCN
    gg = {}
    torch.manual_seed(3); np.random.seed(3)
    with torch.no_grad():
        # (A) GCN alone, feature_dim 128 (4 heads x 32), on the level-2 clouds of the small pyramid (115 + 95 nodes)
        c2 = torch.from_numpy(pyr["points"][2]); l2 = pyr["stack_lengths"][2]
        gcn = GCN(4, 128, 10, ["self", "cross", "self"]).eval()
        f = torch.randn(len(c2), 128)
        c0, c1 = c2[:l2[0]], c2[l2[0]:]
        d0, d1 = gcn(c0.unsqueeze(0).transpose(1, 2), c1.unsqueeze(0).transpose(1, 2),
                     f[:l2[0]].t().unsqueeze(0), f[l2[0]:].t().unsqueeze(0))
        gg["gcn_coords"], gg["gcn_lens"], gg["gcn_feats"] = c2.numpy(), np.asarray(l2, np.int32), f.numpy()
        gg["gcn_out"] = torch.cat([d0, d1], dim=-1).squeeze(0).t().contiguous().numpy()
        for k, v in gcn.state_dict().items():
            gg["gcnsd_" + k] = v.numpy()
        # the kNN lists the reference's get_graph_feature uses (models/gcn.py:48-51), per cloud, global row indices
        from models.gcn import square_distance as sqd
        knn = []
        for cc, off in ((c0, 0), (c1, int(l2[0]))):
            dist = sqd(cc.unsqueeze(0), cc.unsqueeze(0))
            knn.append(dist.topk(k=11, dim=-1, largest=False, sorted=True)[1][0, :, 1:] + off)
        gg["gcn_knn"] = torch.cat(knn).numpy().astype(np.int32)
        # (B) the whole KPFCNN.forward (image_feature False) on the small pair: encoder -> bottle -> GNN -> decoder
        cfg = indoor_cfg(first_feats_dim=32)
        cfg.gnn_feats_dim = 64
        torch.manual_seed(4); np.random.seed(4)
        net = KPFCNN(cfg).eval()
        net.epsilon.data.fill_(-2.0)
        fb = dict(batch)
        fb["features"] = torch.ones(len(pts), 1)
        fb["src_pcd_raw"], fb["tgt_pcd_raw"] = torch.from_numpy(src), torch.from_numpy(tgt)
        fb["stack_lengths"] = [torch.from_numpy(np.asarray(x, np.int32)) for x in pyr["stack_lengths"]]
        res = net(fb)
        gg["net_feats_f"] = res["feats_f"].numpy()
        gg["net_scores_overlap"], gg["net_scores_saliency"] = res["scores_overlap"].numpy(), res["scores_saliency"].numpy()
        for k, v in net.state_dict().items():
            gg["netsd_" + k] = v.numpy()
    np.savez_compressed(f"{OUT}/gnn_ref.npz", **gg)
    print("gnn: gcn out", gg["gcn_out"].shape, "net feats", gg["net_feats_f"].shape)

    # ---- point2node / node visibility (datasets/dataloader.py:70-198) ------------------
    # datasets/dataloader.py cannot be imported (it loads the cp37 cpp_wrappers binaries): the three pure-torch
    # functions are taken from its SOURCE with ast and executed unchanged.
    import ast
    src_text = open(f"{REF}/datasets/dataloader.py").read()
    tree = ast.parse(src_text)
    wanted = ("square_distance", "point2node", "point2node_correspondences")
    ns = {"torch": torch, "np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in wanted:
            exec(compile(ast.Module(body=[node], type_ignores=[]), "dataloader.py", "exec"), ns)
    gn = {}
    rng = np.random.default_rng(7)
    s_nodes, t_nodes = pyr["points"][3][:pyr["stack_lengths"][3][0]], pyr["points"][3][pyr["stack_lengths"][3][0]:]
    corr = np.stack([rng.integers(0, len(src), 900), rng.integers(0, len(tgt), 900)], 1)
    with torch.no_grad():
        sv, tv, si, ti = ns["point2node_correspondences"](torch.from_numpy(s_nodes), torch.from_numpy(src), torch.from_numpy(t_nodes),
                                                           torch.from_numpy(tgt), torch.from_numpy(corr))
    gn.update(src_nodes=s_nodes, tgt_nodes=t_nodes, src_points=src, tgt_points=tgt, corr=corr, src_node_vis=sv.numpy(),
              tgt_node_vis=tv.numpy(), src_idx=si.numpy(), tgt_idx=ti.numpy())
    np.savez_compressed(f"{OUT}/point2node_ref.npz", **gn)
    print("point2node: nodes", len(s_nodes), len(t_nodes), "vis mean", float(sv.mean()), float(tv.mean()))

    # ---- descriptor matching (lib/benchmark_utils.py:76-95, 226-295) ----------------------
    # lib/benchmark_utils.py imports open3d at module level and uses np.bool (removed in NumPy 2): the four pure
    # functions are taken from its SOURCE with ast and executed unchanged, `np` being NumPy plus the old alias np.bool.
    class _NumpyWithOldAliases:
        bool = bool

        def __getattr__(self, k):
            return getattr(np, k)
    bu_tree = ast.parse(open(f"{REF}/lib/benchmark_utils.py").read())
    ns2 = {"torch": torch, "np": _NumpyWithOldAliases()}
    for node in bu_tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("to_tensor", "to_array", "get_inlier_ratio", "mutual_selection"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), "benchmark_utils.py", "exec"), ns2)
    rng = np.random.default_rng(9)
    n_s, n_t = 1500, 1300
    sp_ = rng.uniform(-1, 1, size=(n_s, 3)).astype(np.float32)
    rot_, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    rot_ = (rot_ * np.sign(np.linalg.det(rot_))).astype(np.float32)
    trans_ = rng.normal(size=(3, 1)).astype(np.float32)
    sel = rng.permutation(n_s)[:n_t]
    tp_ = ((rot_ @ sp_.T + trans_).T)[sel] + rng.normal(scale=0.01, size=(n_t, 3)).astype(np.float32)
    fs_ = rng.normal(size=(n_s, 32)).astype(np.float32)
    fs_ /= np.linalg.norm(fs_, axis=1, keepdims=True)
    ft_ = fs_[sel] + rng.normal(scale=0.22, size=(n_t, 32)).astype(np.float32)
    ft_ /= np.linalg.norm(ft_, axis=1, keepdims=True)
    with torch.no_grad():
        res_ = ns2["get_inlier_ratio"](sp_, tp_, fs_, ft_, rot_, trans_)
        scores_ = torch.matmul(torch.from_numpy(fs_), torch.from_numpy(ft_).t())
        mut_ = ns2["mutual_selection"](scores_[None, :, :])[0]
    r_, c_ = np.where(mut_)
    gm = dict(src_pcd=sp_, tgt_pcd=tp_, src_feat=fs_, tgt_feat=ft_, rot=rot_, trans=trans_, mutual_rows=r_.astype(np.int64),
              mutual_cols=c_.astype(np.int64), row_argmax=scores_.max(-1)[1].numpy(),
              inlier_ratio_wo=np.float32(res_["wo"]["inlier_ratio"]), inlier_ratio_w=np.float32(res_["w"]["inlier_ratio"]))
    np.savez_compressed(f"{OUT}/matching_ref.npz", **gm)
    print("matching: mutual pairs", len(r_), "inlier ratios", float(gm["inlier_ratio_wo"]), float(gm["inlier_ratio_w"]))

    # ---- projection --------------------------------------------------------------------------
    from projection import Projection
    gp = {}
    src, tgt, _ = synthetic.match3d_pair(5, n_target=6000)
    views = synthetic.rgbd_views(src, 3, n_views=2, channels=8)
    gp["points"] = src
    x_ref = torch.ones(len(src), 1).repeat(1, 9)
    order = [1, 0]                                  # image 2 written first, image 1 last (architectures.py:367-370)
    res = {}
    for vi, v in enumerate(views):
        pr = Projection(torch.from_numpy(v["intrinsics"]))
        i2, i3 = pr.projection(torch.from_numpy(src), torch.from_numpy(v["depth"])[None], torch.from_numpy(v["world2camera"]))
        res[vi] = (i2, i3)
        for k in ("depth", "world2camera", "intrinsics", "feature2d", "valid_map"):
            gp[f"v{vi}_{k}"] = v[k]
        gp[f"v{vi}_inds2d"], gp[f"v{vi}_inds3d"] = i2.numpy(), i3.numpy()
    for vi in order:
        v = views[vi]
        i2, i3 = res[vi]
        f2d = torch.from_numpy(v["feature2d"]) * torch.from_numpy(v["valid_map"])[None]       # architectures.py:281-285
        feats = f2d[:, i2[:, 1], i2[:, 0]]                                                      # :287
        rows = torch.cat((feats.transpose(1, 0), torch.ones(feats.shape[1], 1)), dim=-1)       # :300
        x_ref[i3, :] = rows                                                                     # :367-368
    gp["x_out"] = x_ref.numpy()
    np.savez_compressed(f"{OUT}/projection_ref.npz", **gp)
    print("projection hits", [len(res[v][1]) for v in res])


if __name__ == "__main__":
    main()
