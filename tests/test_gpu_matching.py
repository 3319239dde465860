"""GPU: descriptor matching front-end and the tester's per-pair .pth file ('next' row 4) vs the NumPy restatement of
lib/benchmark_utils.py (oracle/matching_port.py, pinned to the reference's own functions through tests/golden/matching_ref.npz).  Index lists must
be identical except where two scores tie to within fp32 summation-order noise (reported, bounded)."""
import numpy as np
import pytest
import torch

from oracle import matching_port as mp
from pcrcg_b200 import matching

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = "cuda:0"


def _feats(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1)


@pytest.mark.parametrize("n,m,d", [(1000, 777, 32), (5000, 5000, 32), (64, 3000, 16), (129, 130, 64), (1, 5, 32)])
def test_best_match_and_mutual(n, m, d):
    a, b = _feats(n, d, n), _feats(m, d, m + 1)
    scores = (a.double() @ b.double().t()).numpy()
    idx, val = matching.best_match(a.to(DEV), b.to(DEV), return_scores=True)
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    ref = scores.argmax(1)
    bad = np.nonzero(idx != ref)[0]
    assert len(bad) <= max(1, n // 1000)
    for i in bad:                                            # only fp32-level ties may differ
        assert abs(scores[i, idx[i]] - scores[i, ref[i]]) < 1e-6
    assert np.abs(val - scores[np.arange(n), idx]).max() < 1e-5
    rows, cols = matching.mutual_matches(a.to(DEV), b.to(DEV))
    r_ref, c_ref = mp.mutual_matches(a.numpy(), b.numpy())
    got = set(zip(rows.cpu().tolist(), cols.cpu().tolist()))
    want = set(zip(r_ref.tolist(), c_ref.tolist()))
    assert len(got ^ want) <= max(1, n // 500)
    assert np.all(np.diff(rows.cpu().numpy()) > 0)


def test_first_maximum_rule():
    a = torch.tensor([[1.0] + [0.0] * 31])
    b = torch.zeros(10, 32)
    b[3, 0] = b[7, 0] = 2.0
    assert int(matching.best_match(a.to(DEV), b.to(DEV))[0]) == 3


def test_inlier_ratio_and_pair_file(tmp_path):
    rng = np.random.default_rng(0)
    n = 1500
    src = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
    rot, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    rot = (rot * np.sign(np.linalg.det(rot))).astype(np.float32)
    trans = rng.normal(size=(3, 1)).astype(np.float32)
    perm = rng.permutation(n)
    tgt = ((rot @ src.T + trans).T)[perm] + rng.normal(scale=0.01, size=(n, 3)).astype(np.float32)
    f = _feats(n, 32, 3).numpy()
    ft = f[perm] + rng.normal(scale=0.05, size=(n, 32)).astype(np.float32)
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(DEV)
    res = matching.inlier_ratio(d(src), d(tgt), d(f), d(ft), d(rot), d(trans))
    wo, w = mp.inlier_ratios(src, tgt, f, ft, rot, trans)
    assert abs(float(res["wo"]["inlier_ratio"]) - wo) < 2e-3 and abs(float(res["w"]["inlier_ratio"]) - w) < 2e-3
    assert wo > 0.5
    # the tester's per-pair file (lib/tester.py:92-102)
    path = str(tmp_path / "0.pth")
    pcd = torch.cat([d(src), d(tgt)])
    feats = torch.cat([d(f), d(ft)])
    ov, sa = torch.rand(2 * n, device=DEV), torch.rand(2 * n, device=DEV)
    matching.save_pair(path, pcd, feats, ov, sa, n, d(rot), d(trans))
    back = torch.load(path, weights_only=False)
    assert sorted(back.keys()) == sorted(matching.PAIR_KEYS) and back["len_src"] == n
    assert all(not back[k].is_cuda for k in ("pcd", "feats", "overlaps", "saliency", "rot", "trans"))
    assert torch.equal(back["feats"], feats.cpu()) and back["pcd"].shape == (2 * n, 3)
    assert matching.load_pair(path)["len_src"] == n


def test_matching_vs_reference_golden():
    """mutual pairs, row-wise best matches and both inlier ratios vs the outputs of the reference's own get_inlier_ratio /
    mutual_selection (tests/golden/matching_ref.npz)"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "matching_ref.npz"))
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(DEV)
    idx = matching.best_match(d(g["src_feat"]), d(g["tgt_feat"])).cpu().numpy()
    assert (idx != g["row_argmax"]).sum() <= 1                      # an fp32 summation-order tie at most
    rows, cols = matching.mutual_matches(d(g["src_feat"]), d(g["tgt_feat"]))
    got = set(zip(rows.cpu().tolist(), cols.cpu().tolist()))
    want = set(zip(g["mutual_rows"].tolist(), g["mutual_cols"].tolist()))
    assert len(got ^ want) <= 2
    res = matching.inlier_ratio(d(g["src_pcd"]), d(g["tgt_pcd"]), d(g["src_feat"]), d(g["tgt_feat"]), d(g["rot"]), d(g["trans"]))
    assert abs(float(res["wo"]["inlier_ratio"]) - float(g["inlier_ratio_wo"])) < 2e-3
    assert abs(float(res["w"]["inlier_ratio"]) - float(g["inlier_ratio_w"])) < 2e-3
