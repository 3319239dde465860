"""Descriptor matching front-end and the per-pair result file of the reference's tester ("next" row 4 of the scope
table): what sits between the network output and RANSAC / the 3DMatch evaluation scripts.

  * :func:`best_match`, :func:`mutual_matches`, :func:`inlier_ratio` -- ``lib/benchmark_utils.py:187-224, 226-268, 270-295``
    (``scores = src_feat @ tgt_feat^T`` then argmax / mutual_selection) without materialising the score matrix;
  * :func:`save_pair` / :func:`load_pair` -- the ``.pth`` dict ``lib/tester.py:92-102`` writes per pair
    (``pcd, feats, overlaps, saliency, len_src, rot, trans``), readable by the reference's ``scripts/evaluate_predator.py``.
RANSAC itself (open3d) stays with the reference.
"""
import numpy as np
import torch

from . import ops
from ._lib import lib, check


def best_match(a, b, return_scores=False):
    """a [n,D], b [m,D] CUDA fp32 -> int32 [n]: argmax_j <a_i, b_j> (first maximum)."""
    ops._need_cuda(a, b)
    a, b = ops._f32c(a), ops._f32c(b)
    n, d = a.shape
    if b.shape[1] != d:
        raise RuntimeError("best_match: descriptor lengths differ")
    idx = torch.empty(n, dtype=torch.int32, device=a.device)
    val = torch.empty(n, dtype=torch.float32, device=a.device) if return_scores else None
    with torch.cuda.device(a.device):
        check(lib().pcrcg_best_match_dev(a.data_ptr(), n, b.data_ptr(), b.shape[0], d, idx.data_ptr(),
                                         val.data_ptr() if val is not None else None, ops._stream()))
    return (idx, val) if return_scores else idx


def mutual_matches(src_feat, tgt_feat):
    """-> (row_sel, col_sel) int64 device tensors == np.where(mutual_selection(src_feat @ tgt_feat^T))
    (lib/benchmark_utils.py:199-201, 270-295): pairs that are each other's best match, ascending in row."""
    row_best = best_match(src_feat, tgt_feat)
    col_best = best_match(tgt_feat, src_feat)
    n = row_best.shape[0]
    mutual = torch.empty(n, dtype=torch.uint8, device=row_best.device)
    with torch.cuda.device(row_best.device):
        check(lib().pcrcg_mutual_dev(row_best.data_ptr(), col_best.data_ptr(), n, mutual.data_ptr(), ops._stream()))
    rows = torch.nonzero(mutual, as_tuple=False).view(-1)
    return rows, row_best[rows].long()


def inlier_ratio(src_pcd, tgt_pcd, src_feat, tgt_feat, rot, trans, inlier_distance_threshold=0.1):
    """get_inlier_ratio (lib/benchmark_utils.py:226-268): {'wo': ..., 'w': ...} inlier ratios without / with mutual check."""
    src = (rot @ src_pcd.t() + trans).t()
    idx = best_match(src_feat, tgt_feat).long()
    d_wo = torch.norm(src - tgt_pcd[idx], dim=1)
    rows, cols = mutual_matches(src_feat, tgt_feat)
    d_w = torch.norm(src[rows] - tgt_pcd[cols], dim=1)
    f = lambda d: (d < inlier_distance_threshold).float().mean()
    return {"wo": {"distance": d_wo.cpu().numpy(), "inlier_ratio": f(d_wo).cpu()},
            "w": {"distance": d_w.cpu().numpy(), "inlier_ratio": f(d_w).cpu()}}


PAIR_KEYS = ("pcd", "feats", "overlaps", "saliency", "len_src", "rot", "trans")


def save_pair(path, pcd, feats, overlaps, saliency, len_src, rot, trans):
    """lib/tester.py:92-102: one ``<idx>.pth`` per pair, CPU tensors, same keys and dtypes."""
    cpu = lambda t: t.detach().cpu() if torch.is_tensor(t) else torch.as_tensor(np.asarray(t))
    data = dict(pcd=cpu(pcd), feats=cpu(feats), overlaps=cpu(overlaps), saliency=cpu(saliency),
                len_src=int(len_src), rot=cpu(rot), trans=cpu(trans))
    torch.save(data, path)
    return data


def load_pair(path):
    data = torch.load(path, map_location="cpu", weights_only=False)
    missing = [k for k in PAIR_KEYS if k not in data]
    if missing:
        raise RuntimeError(f"{path}: not a PCR-CG pair file (missing {missing})")
    return data
