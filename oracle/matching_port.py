"""TEST INFRASTRUCTURE ONLY: NumPy restatement of the descriptor-matching helpers of the reference
(lib/benchmark_utils.py).  The module itself cannot be imported here (open3d at module level, np.bool removed in NumPy 2);
tests/golden/make_golden.py executes the SOURCE of its four pure functions (to_tensor, to_array, get_inlier_ratio,
mutual_selection) unchanged, with `np` = NumPy plus the old alias, and tests/test_oracle_pinning.py pins this restatement to
those outputs (tests/golden/matching_ref.npz)."""
import numpy as np


def mutual_selection(score_mat):
    """lib/benchmark_utils.py:270-295: 1 where an entry is the maximum of its row AND of its column (first maxima)."""
    flag_row = np.zeros_like(score_mat)
    flag_col = np.zeros_like(score_mat)
    np.put_along_axis(flag_row, np.argmax(score_mat, 1)[:, None], 1, 1)
    np.put_along_axis(flag_col, np.argmax(score_mat, 0)[None, :], 1, 0)
    return flag_row.astype(bool) & flag_col.astype(bool)


def mutual_matches(src_feat, tgt_feat):
    """lib/benchmark_utils.py:199-201: np.where(mutual_selection(src @ tgt^T))"""
    scores = src_feat.astype(np.float32) @ tgt_feat.astype(np.float32).T
    return np.where(mutual_selection(scores))


def inlier_ratios(src_pcd, tgt_pcd, src_feat, tgt_feat, rot, trans, thr=0.1):
    """lib/benchmark_utils.py:226-268"""
    src = (rot @ src_pcd.T + trans).T
    scores = src_feat @ tgt_feat.T
    idx = scores.argmax(-1)
    wo = float((np.linalg.norm(src - tgt_pcd[idx], axis=1) < thr).mean())
    r, c = np.where(mutual_selection(scores))
    w = float((np.linalg.norm(src[r] - tgt_pcd[c], axis=1) < thr).mean())
    return wo, w
