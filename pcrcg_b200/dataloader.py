"""Pyramid construction with the reference's function names (``datasets/dataloader.py``), running
on the GPU through libpcrcg_b200.so.

  batch_grid_subsampling_kpconv  <- datasets/dataloader.py:14-52
  batch_neighbors_kpconv         <- datasets/dataloader.py:54-69
  collate_fn_descriptor          <- datasets/dataloader.py:203-400 (the pyramid loop :239-359)
  calibrate_neighbors            <- datasets/dataloader.py:402-434

Where the call happens differs from the reference on purpose: the reference builds the pyramid in
forked DataLoader workers on the CPU; CUDA cannot be initialised there, so these run in the main
process on the device that will also run the encoder (the tensors never leave HBM).
"""
import numpy as np
import torch

from . import ops


def _dev_f32(t, device):
    if not torch.is_tensor(t):
        t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32))
    return t.to(device=device, dtype=torch.float32, non_blocking=True)


def _dev_i32(t, device):
    if not torch.is_tensor(t):
        t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.int32))
    return t.to(device=device, dtype=torch.int32, non_blocking=True)


def _device_of(*ts):
    """the CUDA device of the first device tensor among the arguments, else the current CUDA device: the reference hands these
    functions CPU tensors / arrays (datasets/dataloader.py:273-301); here they are copied to the GPU, where all the work happens"""
    for t in ts:
        if torch.is_tensor(t) and t.is_cuda:
            return t.device
    return torch.device("cuda")


def batch_grid_subsampling_kpconv(points, batches_len, features=None, labels=None, sampleDl=0.1, max_p=0, verbose=0,
                                  random_grid_orient=True):
    """-> (s_points [M,3] f32, s_len [B] i32[, s_features][, s_labels]) in the reference's order (datasets/dataloader.py:18-52),
    on the device of ``points`` (cuda)."""
    dev = _device_of(points, batches_len)
    points, batches_len = _dev_f32(points, dev), _dev_i32(batches_len, dev)
    if features is None and labels is None:
        return ops.subsample_batch(points, batches_len, sampleDl, max_p)
    features = None if features is None else _dev_f32(features, dev)
    labels = None if labels is None else _dev_i32(labels, dev)
    return ops.subsample_batch_ex(points, batches_len, sampleDl, max_p, features=features, classes=labels)


def batch_neighbors_kpconv(queries, supports, q_batches, s_batches, radius, max_neighbors):
    """-> int32 [Nq, min(max_neighbors, max_count)] (max_neighbors <= 0: full width), on the device."""
    dev = _device_of(queries, supports)
    return ops.batch_query(_dev_f32(queries, dev), _dev_f32(supports, dev), _dev_i32(q_batches, dev), _dev_i32(s_batches, dev), radius,
                           max_neighbors)


@torch.no_grad()
def build_pyramid(points, lengths, config, neighborhood_limits, device=None, pairs=True, long_indices=False,
                  return_counts=False):
    """The pyramid loop of ``collate_fn_descriptor`` (datasets/dataloader.py:239-359).

    points [N0,3], lengths [B] = the stacked clouds of one OR MANY fragment pairs
    ([src0, tgt0, src1, tgt1, ...]); returns the reference's dict keys ``points, neighbors, pools,
    upsamples, stack_lengths`` (lists per layer, device tensors; index lists int32 unless
    ``long_indices``) plus ``pair_segments`` (int32 row starts of each pair, per layer) used to keep
    InstanceNorm statistics per pair.  Index-list widths follow the reference:
    min(limit, max_count) (datasets/dataloader.py:66-69)."""
    device = torch.device(device if device is not None else (points.device if torch.is_tensor(points) and points.is_cuda else "cuda"))
    pts = _dev_f32(points, device)
    lens = _dev_i32(lengths, device)
    arch = config.architecture
    r_normal = config.first_subsampling_dl * config.conv_radius
    out = dict(points=[], neighbors=[], pools=[], upsamples=[], stack_lengths=[])
    pending, counts_out = [], []            # (list name, layer, rows, limit, max_count tensor)
    layer, layer_blocks = 0, []
    next_grid = None          # the upsample grid of level l (coarse points, radius 2r) is the conv/pool grid of level l+1
    for block_i, block in enumerate(arch):
        if "global" in block or "upsample" in block:
            break
        if not ("pool" in block or "strided" in block):
            layer_blocks.append(block)
            if block_i < len(arch) - 1 and "upsample" not in arch[block_i + 1]:
                continue
        limit = int(neighborhood_limits[layer])
        grid_fine = next_grid
        next_grid = None
        if layer_blocks:
            if grid_fine is None:
                grid_fine = ops.RadiusGrid(pts, lens, r_normal)
            conv_i, cnt, mx = grid_fine.query(pts, lens, limit, want_counts=return_counts)
            pending.append(("neighbors", layer, conv_i, limit, mx))
            counts_out.append(cnt)
        else:
            conv_i = torch.zeros((0, 1), dtype=torch.int32, device=device)
        if "pool" in block or "strided" in block:
            dl = 2 * r_normal / config.conv_radius
            pool_p, pool_b = ops.subsample_batch(pts, lens, dl)
            if grid_fine is None:
                grid_fine = ops.RadiusGrid(pts, lens, r_normal)
            pool_i, _, mxp = grid_fine.query(pool_p, pool_b, limit, want_counts=False)
            pending.append(("pools", layer, pool_i, limit, mxp))
            next_grid = ops.RadiusGrid(pool_p, pool_b, 2 * r_normal)
            up_i, _, mxu = next_grid.query(pts, lens, limit, want_counts=False)
            pending.append(("upsamples", layer, up_i, limit, mxu))
        else:
            pool_i = torch.zeros((0, 1), dtype=torch.int32, device=device)
            pool_p = torch.zeros((0, 3), dtype=torch.float32, device=device)
            pool_b = torch.zeros((0,), dtype=torch.int32, device=device)
            up_i = torch.zeros((0, 1), dtype=torch.int32, device=device)
        out["points"].append(pts)
        out["neighbors"].append(conv_i)
        out["pools"].append(pool_i)
        out["upsamples"].append(up_i)
        out["stack_lengths"].append(lens)
        pts, lens = pool_p, pool_b
        r_normal *= 2
        layer += 1
        layer_blocks = []
    # one host sync for all widths: width = min(limit, max_count)
    if pending:
        mxs = torch.cat([p[4] for p in pending]).cpu().tolist()
        for (name, l, rows, limit, _), mx in zip(pending, mxs):
            w = min(limit, int(mx))
            out[name][l] = rows if w == limit else rows[:, :w]
    if long_indices:
        for name in ("neighbors", "pools", "upsamples"):
            out[name] = [t.long() for t in out[name]]
    # per-pair row starts at every layer
    segs = []
    for l in out["stack_lengths"]:
        paired = pairs and l.numel() % 2 == 0 and l.numel() > 0
        segs.append(ops.group_starts(l, 2 if paired else max(1, l.numel())))
    out["pair_segments"] = segs
    if return_counts:
        out["neighbor_counts"] = counts_out
    return out


def collate_fn_descriptor(list_data, config, neighborhood_limits, device="cuda"):
    """datasets/dataloader.py:203-400.  ``list_data`` items are dicts with ``src_pcd``, ``tgt_pcd`` ([N,3]) and ``src_feats``,
    ``tgt_feats`` ([N,C]); unlike the reference any number of pairs may be stacked.  The pyramid keys (``points, neighbors, pools,
    upsamples, stack_lengths, features``) are device tensors; for a single pair (the reference's only case, :207) the dict also
    carries the reference's other keys, see :func:`_reference_extras`."""
    pts, lens, feats = [], [], []
    for d in list_data:
        for k in ("src", "tgt"):
            p = np.asarray(d[f"{k}_pcd"], dtype=np.float32)
            pts.append(p)
            lens.append(len(p))
            feats.append(np.asarray(d[f"{k}_feats"], dtype=np.float32))
    batch = build_pyramid(np.concatenate(pts), np.array(lens, np.int32), config, neighborhood_limits, device=device)
    batch["features"] = _dev_f32(np.concatenate(feats), batch["points"][0].device)
    if len(list_data) == 1:
        _reference_extras(batch, list_data[0], pts[0], pts[1])
    return batch


# keys the reference's collate copies from the dataset item unchanged (datasets/dataloader.py:380-397)
_IMAGE_KEYS = ("src1_inds2d", "src2_inds2d", "src3_inds2d", "tgt1_inds2d", "tgt2_inds2d", "tgt3_inds2d", "src1_inds3d", "src2_inds3d",
               "src3_inds3d", "tgt1_inds3d", "tgt2_inds3d", "tgt3_inds3d", "src_color1", "src_color2", "src_color3", "tgt_color1",
               "tgt_color2", "tgt_color3", "id_name", "detect_1", "detect_2", "detect_3", "detect_4", "des1", "des2", "des3", "des4",
               "src_valid_map1", "src_valid_map2", "tgt_valid_map1", "tgt_valid_map2")


def _reference_extras(batch, item, src, tgt):
    """The remaining keys of the reference's dict for ONE pair (datasets/dataloader.py:359-397), as far as the dataset item carries
    their inputs: rot / trans / correspondences / sample and the image keys are handed through, the raw clouds are returned as
    float tensors, and the node labels of the coarsest level (``node_overlap_gt``, ``points2node``, :309-322) are computed on the
    device from the ground-truth correspondences."""
    for k in ("rot", "trans"):
        if k in item:
            batch[k] = torch.as_tensor(np.asarray(item[k]))
    batch["src_pcd_raw"] = torch.from_numpy(np.ascontiguousarray(src)).float()
    batch["tgt_pcd_raw"] = torch.from_numpy(np.ascontiguousarray(tgt)).float()
    if "sample" in item:
        batch["sample"] = item["sample"]
    if "correspondences" in item:
        batch["correspondences"] = item["correspondences"]
        nodes = batch["points"][-1]
        n_src = int(batch["stack_lengths"][-1][0].item())
        sv, tv, s2n, t2n = point2node_correspondences(nodes[:n_src], src, nodes[n_src:], tgt, item["correspondences"])
        batch["node_overlap_gt"] = torch.cat((sv, tv))
        batch["points2node"] = torch.cat((s2n, t2n))
    for key in item:
        if key in _IMAGE_KEYS:
            batch[key] = item[key]


@torch.no_grad()
def _calibration_pairs(dataset):
    """the reference walks ``dataset[i] for i in range(len(dataset))`` (datasets/dataloader.py:411-413); any iterable works here.
    Items: the dataset's dicts (``src_pcd`` / ``tgt_pcd``) or plain (src, tgt) tuples."""
    if hasattr(dataset, "__len__") and hasattr(dataset, "__getitem__") and not isinstance(dataset, (list, tuple)):
        items = (dataset[i] for i in range(len(dataset)))
    else:
        items = iter(dataset)
    for it in items:
        yield (it["src_pcd"], it["tgt_pcd"]) if isinstance(it, dict) else (it[0], it[1])


def calibrate_neighbors(dataset, config, collate_fn=None, keep_ratio=0.8, samples_threshold=2000, device="cuda"):
    """datasets/dataloader.py:402-434, same signature: histogram of neighbourhood sizes per layer over the pairs of ``dataset``,
    limit = number of histogram bins whose cumulated mass stays below keep_ratio.  ``collate_fn`` is accepted for call
    compatibility and not used: the un-truncated neighbour counts come straight from the search kernel instead of from
    905-column index lists."""
    hist_n = int(np.ceil(4 / 3 * np.pi * (config.deform_radius + 1) ** 3))
    neighb_hists = np.zeros((config.num_layers, hist_n), dtype=np.int64)
    for src, tgt in _calibration_pairs(dataset):
        pts = np.concatenate([src, tgt]).astype(np.float32)
        lens = np.array([len(src), len(tgt)], np.int32)
        b = build_pyramid(pts, lens, config, [hist_n] * 5, device=device, return_counts=True)
        counts = [np.minimum(c.cpu().numpy(), hist_n) for c in b["neighbor_counts"]]
        hists = [np.bincount(c, minlength=hist_n)[:hist_n] for c in counts]
        neighb_hists += np.vstack(hists)
        if np.min(np.sum(neighb_hists, axis=1)) > samples_threshold:
            break
    cumsum = np.cumsum(neighb_hists.T, axis=0)
    return np.sum(cumsum < (keep_ratio * cumsum[hist_n - 1, :]), axis=0)


# ---- node labels of the collate (datasets/dataloader.py:91-198, 309-322) -----------------------------------------------
def _p2n(nodes, node_lens, points, point_lens):
    dev = points.device
    ps, ns = ops.cloud_starts(point_lens).to(dev), ops.cloud_starts(node_lens).to(dev)
    out = torch.empty(points.shape[0], dtype=torch.int32, device=dev)
    from ._lib import lib, check
    with torch.cuda.device(dev):
        check(lib().pcrcg_point2node_dev(points.data_ptr(), points.shape[0], ps.data_ptr(), nodes.data_ptr(), ns.data_ptr(),
                                         ns.shape[0] - 1, out.data_ptr(), ops._stream()))
    return out, ps, ns


def point2node(nodes, points):
    """datasets/dataloader.py:91-106: index [N] (int64) of the nearest node of every point (one cloud)."""
    nodes = _dev_f32(nodes, nodes.device if torch.is_tensor(nodes) and nodes.is_cuda else "cuda")
    points = _dev_f32(points, nodes.device)
    out, _, _ = _p2n(nodes.contiguous(), [nodes.shape[0]], points.contiguous(), [points.shape[0]])
    return out.long()


def point2node_correspondences(src_nodes, src_points, tgt_nodes, tgt_points, point_correspondences, device=None):
    """datasets/dataloader.py:108-198 -> (src_node_vis, tgt_node_vis, src_idx, tgt_idx): per node the fraction of its
    points that have a ground-truth correspondence (nodes without points: 0 / 1), and the point -> node assignment."""
    dev = src_nodes.device if torch.is_tensor(src_nodes) and src_nodes.is_cuda else torch.device("cuda")
    sn, sp, tn, tp = (_dev_f32(t, dev).contiguous() for t in (src_nodes, src_points, tgt_nodes, tgt_points))
    corr = torch.as_tensor(point_correspondences).to(dev).long()
    nodes, pts = torch.cat([sn, tn]), torch.cat([sp, tp])
    p2n, ps, ns = _p2n(nodes, [sn.shape[0], tn.shape[0]], pts, [sp.shape[0], tp.shape[0]])
    visible = torch.zeros(pts.shape[0], dtype=torch.uint8, device=dev)
    visible[corr[:, 0]] = 1
    visible[corr[:, 1] + sp.shape[0]] = 1
    tot = torch.zeros(nodes.shape[0], dtype=torch.int32, device=dev)
    vis = torch.zeros(nodes.shape[0], dtype=torch.int32, device=dev)
    from ._lib import lib, check
    with torch.cuda.device(dev):
        check(lib().pcrcg_node_counts_dev(p2n.data_ptr(), visible.data_ptr(), pts.shape[0], ps.data_ptr(), ns.data_ptr(), 2,
                                          tot.data_ptr(), vis.data_ptr(), ops._stream()))
    ratio = vis.float() / tot.clamp_min(1).float()                     # src_tot_num defaults to 1 (:135,148)
    n_s, n_p = sn.shape[0], sp.shape[0]
    return ratio[:n_s], ratio[n_s:], p2n[:n_p].long(), p2n[n_p:].long()
