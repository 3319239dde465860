"""GPU parity of the ONE-kernel KPConv (csrc/kpconv_fused.cu: gather -> influence -> tcgen05 contraction with the weights
resident in tensor memory) against the PyTorch-fp32 restatement of models/blocks.py:229-374 (oracle/blocks_port.py, pinned
to the reference's goldens by tests/test_oracle_pinning.py) and against the two-kernel path, for every channel count of the
backbone (64 / 128 / 256 / 512), int32 and int64 lists, conv and strided (pool) geometry, ragged sizes and several
InstanceNorm segments.  Tolerance (north_star): 1e-3 normwise; the measured error is ~1e-5 and asserted at 1e-4."""
import numpy as np
import pytest
import torch

from oracle import blocks_port as bp
from pcrcg_b200 import ops, synthetic
from pcrcg_b200._lib import lib, check

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _d(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def _err(a, ref):
    a, ref = a.detach().cpu().double(), ref.detach().cpu().double()
    return float((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def _planes(x):
    """bf16 (hi, lo) planes + (row sum > 0) flags of fp32 features, as instance_norm_act(planes_only=True) emits them"""
    n, c = x.shape
    hi, lo, ld = ops._split_planes(n, c, x.device)
    check(lib().pcrcg_split_bf16_dev(x.data_ptr(), c, n, c, hi.data_ptr(), lo.data_ptr(), ld, ops._stream()))
    return ops.PlaneTensor(hi, lo, ld, n, c, (x.sum(1) > 0).to(torch.uint8))


@pytest.fixture(autouse=True)
def _restore_option():
    yield
    check(lib().pcrcg_set_option(b"kpconv_fused", 1))


def _geometry(seed, n_target, limit, n_pairs=2, strided=False):
    pairs = [synthetic.match3d_pair(seed + k, n_target=n_target + 37 * k)[:2] for k in range(n_pairs)]
    pts = np.concatenate([np.concatenate(p) for p in pairs]).astype(np.float32)
    lens = np.array([len(c) for p in pairs for c in p], np.int32)
    P, L = _d(pts), _d(lens)
    if strided:
        q, ql = ops.subsample_batch(P, L, 0.05)
        rows = ops.batch_query(q, P, ql, L, 0.0625, limit)
    else:
        q, ql = P, L
        rows = ops.batch_query(P, P, L, L, 0.0625, limit)
    seg = torch.cat([torch.zeros(1, dtype=torch.int32, device=DEV), ql.view(-1, 2).sum(1).cumsum(0).to(torch.int32)])
    return q, P, rows, seg


def _run(mode, q, s, rows, xp, kp, w, seg):
    check(lib().pcrcg_set_option(b"kpconv_fused", mode))
    out = ops.kpconv_forward(q, s, rows, xp, kp, w, 0.05, stat_segments=seg)
    mean, rstd, _, _ = ops.attached(out, "_pcrcg_stats")
    torch.cuda.synchronize()
    return out, mean, rstd


@pytest.mark.parametrize("cin,cout", [(64, 64), (128, 128), (256, 256), (512, 512), (64, 128), (128, 64)])
@pytest.mark.parametrize("idx_dtype", [torch.int32, torch.int64])
def test_fused_vs_port_and_two_kernel(cin, cout, idx_dtype):
    q, s, rows, seg = _geometry(3, 1100 if cin <= 128 else 500, 34)
    rows = rows.to(idx_dtype)
    g = torch.Generator().manual_seed(cin + cout)
    x = torch.randn(s.shape[0], cin, generator=g)
    x[::7] = -x[::7].abs()                              # rows with a non-positive sum: the neighbour count excludes them
    w = torch.randn(15, cin, cout, generator=g) / np.sqrt(15 * cin)
    kp = torch.randn(15, 3, generator=g) * 0.03
    ref = bp.kpconv(q.cpu(), s.cpu(), rows.cpu().long(), x, kp, w, 0.05)
    xp = _planes(x.to(DEV))
    fused, fm, fr = _run(2, q, s, rows, xp, kp.to(DEV), w.to(DEV), seg)
    two, tm, tr = _run(0, q, s, rows, xp, kp.to(DEV), w.to(DEV), seg)
    assert _err(fused, ref) < 1e-4, _err(fused, ref)
    assert _err(two, ref) < 1e-4
    # InstanceNorm statistics of the result (per pair): same as the two-kernel epilogue's
    assert _err(fm, tm) < 1e-4 and _err(fr, tr) < 1e-4
    # and equal to a direct computation on the fused output
    seg_h = seg.cpu().tolist()
    for k in range(len(seg_h) - 1):
        blk = fused[seg_h[k]:seg_h[k + 1]].double()
        assert torch.allclose(blk.mean(0).float(), fm[k], atol=2e-6 + 1e-5 * float(blk.abs().max()))
        assert _err(fr[k], (1.0 / torch.sqrt(blk.var(0, unbiased=False) + 1e-5)).float()) < 1e-4


@pytest.mark.parametrize("limit", [20, 50, 64])
def test_fused_strided_geometry_and_wide_lists(limit):
    """pool lists (Nq != Ns), list widths up to the kernel's 64, a shadow-heavy tail"""
    q, s, rows, seg = _geometry(11, 900, limit, n_pairs=3, strided=True)
    g = torch.Generator().manual_seed(limit)
    x = torch.randn(s.shape[0], 64, generator=g)
    w = torch.randn(15, 64, 64, generator=g) / np.sqrt(15 * 64)
    kp = torch.randn(15, 3, generator=g) * 0.03
    ref = bp.kpconv(q.cpu(), s.cpu(), rows.cpu().long(), x, kp, w, 0.05)
    fused, _, _ = _run(1, q, s, rows, _planes(x.to(DEV)), kp.to(DEV), w.to(DEV), seg)
    assert fused.shape == ref.shape and _err(fused, ref) < 1e-4


def test_fused_tiny_and_empty_neighbourhoods():
    """fewer points than one 8-point tile; queries whose lists hold shadows only give exact zeros"""
    g = torch.Generator().manual_seed(5)
    s = torch.rand(37, 3, generator=g) * 0.1
    q = torch.cat([s[:5], torch.full((2, 3), 9.0)])                       # the last two queries have no neighbour
    lens_s, lens_q = torch.tensor([37], dtype=torch.int32), torch.tensor([7], dtype=torch.int32)
    rows = ops.batch_query(q.to(DEV), s.to(DEV), lens_q.to(DEV), lens_s.to(DEV), 0.05, 0)
    rows = torch.cat([rows[:5], torch.full((2, rows.shape[1]), 37, dtype=torch.int32, device=DEV)])
    x = torch.randn(37, 64, generator=g)
    w = torch.randn(15, 64, 64, generator=g) / 31.0
    kp = torch.randn(15, 3, generator=g) * 0.02
    ref = bp.kpconv(q, s, rows.cpu().long(), x, kp, w, 0.04)
    check(lib().pcrcg_set_option(b"kpconv_fused", 1))
    out = ops.kpconv_forward(q.to(DEV), s.to(DEV), rows, _planes(x.to(DEV)), kp.to(DEV), w.to(DEV), 0.04, stat_segments=True)
    assert _err(out, ref) < 1e-4
    assert float(out[5:].abs().max()) == 0.0


def test_fused_is_the_default_for_64_channels():
    """the bottleneck KPConv of the first two encoder stages (64 -> 64) takes the one-kernel path by default"""
    q, s, rows, seg = _geometry(1, 800, 30)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(s.shape[0], 64, generator=g).to(DEV)
    w = (torch.randn(15, 64, 64, generator=g) / 31.0).to(DEV)
    kp = (torch.randn(15, 3, generator=g) * 0.03).to(DEV)
    L = lib()
    L.pcrcg_profile_enable(1)
    try:
        ops.kpconv_forward(q, s, rows, _planes(x), kp, w, 0.05, stat_segments=seg)
        names = [L.pcrcg_profile_class_name(i).decode() for i in range(L.pcrcg_profile_classes())]
        import ctypes
        ms = (ctypes.c_double * len(names))()
        cnt = (ctypes.c_int64 * len(names))()
        check(L.pcrcg_profile_report(ms, cnt))
    finally:
        L.pcrcg_profile_enable(0)
    t = dict(zip(names, ms))
    assert t["kpconv_fused"] > 0.0 and t["kpconv_aggregate"] == 0.0
