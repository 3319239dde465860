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
