"""GPU: the dense contraction (tcgen05 bf16x3 tensor-core kernel and the fp32 CUDA-core kernel)
vs a float64 torch matmul.  Tolerance: normwise max|a-b|/max|ref| (features tolerance is 1e-3; the
contraction itself must be far inside it: <= 5e-5 tensor path, <= 1e-5 CUDA-core path)."""
import pytest
import torch

from pcrcg_b200 import ops

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]
DEV = "cuda:0"

SHAPES = [  # (M, N, K)
    (128, 64, 64), (128, 128, 128), (300, 64, 960), (1000, 128, 1920), (257, 256, 3840), (130, 512, 7680),
    (777, 16, 48), (512, 32, 200), (4096, 2048, 512), (5000, 64, 256), (64, 1024, 256), (1, 64, 64),
    (333, 64, 15), (200, 24, 240), (100, 128, 1935),      # ragged: CUDA-core fallback
]


def _err(a, ref):
    return float((a.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("path", ["tensor", "simt"])
def test_matmul_kn(M, N, K, path):
    ops.force_simt_contraction(path == "simt")
    try:
        g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
        a = torch.randn(M, K, generator=g).to(DEV)
        b = (torch.randn(K, N, generator=g) / K ** 0.5).to(DEV)
        out = ops.matmul(a, b)
        torch.cuda.synchronize()
        ref = a.double() @ b.double()
        assert out.shape == (M, N)
        assert _err(out, ref) < (5e-5 if path == "tensor" else 1e-5)
    finally:
        ops.force_simt_contraction(False)


@pytest.mark.parametrize("M,N,K", [(1000, 256, 64), (4097, 64, 256), (500, 2048, 512), (900, 128, 129)])
def test_linear_nk(M, N, K):
    g = torch.Generator().manual_seed(N)
    x = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    out = ops.linear(x, w)
    ref = x.double() @ w.double().t()
    assert _err(out, ref) < 5e-5


def test_large_and_repeatable():
    g = torch.Generator().manual_seed(1)
    a = torch.randn(200_000, 960, generator=g).to(DEV)
    b = (torch.randn(960, 64, generator=g) / 31.0).to(DEV)
    o1 = ops.matmul(a, b)
    o2 = ops.matmul(a, b)
    assert torch.equal(o1, o2)
    ref = a[:4096].double() @ b.double()
    assert _err(o1[:4096], ref) < 5e-5
    ref = a[-4096:].double() @ b.double()
    assert _err(o1[-4096:], ref) < 5e-5


def _ref_stats(out, seg, eps=1e-5):
    means, rstds = [], []
    for a, b in zip(seg[:-1], seg[1:]):
        blk = out[a:b].double()
        mu = blk.mean(0)
        var = (blk * blk).mean(0) - mu * mu
        means.append(mu)
        rstds.append(1.0 / torch.sqrt(var.clamp_min(0) + eps))
    return torch.stack(means), torch.stack(rstds)


@pytest.mark.parametrize("M,N,K,seg", [
    (1000, 256, 64, None),                       # one group (the reference's case)
    (4097, 64, 256, [0, 1000, 1031, 4097]),      # boundaries inside a 32-row epilogue warp and inside a tile
    (777, 16, 48, [0, 128, 777]),                # boundary on a tile edge, narrow N
    (5000, 2048, 128, [0, 2500, 5000]),
    (130, 512, 512, [0, 1, 130]),                # a one-row segment
])
def test_linear_epilogue_statistics(M, N, K, seg):
    """InstanceNorm statistics accumulated by the contraction epilogue == statistics of the written output
    (fp64 torch), and instance_norm_act consumes them (same result as the separate statistics pass)."""
    g = torch.Generator().manual_seed(M + N)
    x = (torch.randn(M, K, generator=g) + 0.3).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    hi, lo, ld = ops._split_planes(M, K, x.device)
    from pcrcg_b200._lib import lib, check
    check(lib().pcrcg_split_bf16_dev(x.data_ptr(), K, M, K, hi.data_ptr(), lo.data_ptr(), ld, torch.cuda.current_stream().cuda_stream))
    ops._attach(x, "_pcrcg_split", (hi, lo, ld))
    segt = None if seg is None else torch.tensor(seg, dtype=torch.int32, device=DEV)
    out = ops.linear(x, w, stat_segments=True if segt is None else segt)
    assert hasattr(out, "_pcrcg_stats"), "tensor-core path must attach the statistics"
    mean, rstd, _, _ = ops.attached(out, "_pcrcg_stats")
    rm, rr = _ref_stats(out, seg or [0, M])
    assert float((mean.double() - rm).abs().max()) < 1e-5 * float(rm.abs().max().clamp_min(1.0))
    # variance = E[x^2] - mean^2 with fp32 partial sums per 32-row block (as the separate statistics pass): absolute
    # error ~1e-7 E[x^2]; compared as variances so that a degenerate one-row segment (var = 0, rstd = eps^-1/2) is judged fairly
    var, var_ref = 1.0 / rstd.double() ** 2 - 1e-5, 1.0 / rr ** 2 - 1e-5
    assert float(((var - var_ref).abs() / (var_ref + rm * rm + 1e-5)).max()) < 2e-6
    fused = ops.instance_norm_act(out, segt, 0.1)
    plain = out.clone()                          # no attribute -> separate statistics pass
    sep = ops.instance_norm_act(plain, segt, 0.1)
    assert float((fused - sep).abs().max()) < 2e-4 * float(sep.abs().max())


def test_derived_data_is_dropped_after_an_in_place_update():
    """bf16 planes / statistics ride on tensors as attributes stamped with the tensor's version: an in-place update of the
    values (torch's add_, or this module's own in-place bias_act) must not leave stale planes behind"""
    g = torch.Generator().manual_seed(7)
    x = torch.randn(300, 64, generator=g).to(DEV)
    w = (torch.randn(32, 64, generator=g) / 8.0).to(DEV)
    y = ops.instance_norm_act(x, None, 0.1, emit_split=True)
    assert ops.attached(y, "_pcrcg_split") is not None
    ref0 = ops.linear(y.clone(), w)
    assert float((ops.linear(y, w) - ref0).abs().max()) < 1e-4 * float(ref0.abs().max())
    y.add_(1.0)                                            # version bump: the planes describe the old values
    assert ops.attached(y, "_pcrcg_split") is None
    ref1 = ops.linear(y.clone(), w)
    assert float((ops.linear(y, w) - ref1).abs().max()) < 1e-4 * float(ref1.abs().max())
    z = ops.instance_norm_act(x, None, 0.1, emit_split=True)
    ops.bias_act(z, torch.ones(64, device=DEV), out=z)     # in place through the library: derived data stripped
    assert not hasattr(z, "_pcrcg_split")
    ref2 = ops.linear(z.clone(), w)
    assert float((ops.linear(z, w) - ref2).abs().max()) < 1e-4 * float(ref2.abs().max())
