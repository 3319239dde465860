#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 bf16x3 contraction on the hot path's shapes (16 stacked pairs).
Prints per shape: device time of the core kernel (operands pre-split), useful TFLOP/s, tensor-pipe
TFLOP/s (x3), and the HBM floor of reading A + writing C."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pcrcg_b200._lib import lib, check  # noqa: E402

SHAPES = [("conv L0 64->64", 646588, 64, 960), ("unary1 L0 128->64", 646588, 64, 128), ("unary2 L0 64->256", 646588, 256, 64),
          ("shortcut L0 128->256", 646588, 256, 128), ("unary1 L0 256->64", 646588, 64, 256), ("conv L1 128->128", 156404, 128, 1920),
          ("unary L1 256->512", 156404, 512, 256), ("conv L2 256->256", 42419, 256, 3840), ("unary L2 512->1024", 42419, 1024, 512),
          ("conv L3 512->512", 11538, 512, 7680), ("unary L3 1024->2048", 11538, 2048, 1024)]


def main():
    L = lib()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    for a in sys.argv:
        if a.startswith("--dbg="):
            check(L.pcrcg_set_option(b"stats_debug", int(a[6:])))
    for name, M, N, K in SHAPES:
        if "--wide" in sys.argv and N < 256:
            continue
        ldk = (K + 7) // 8 * 8
        a = torch.randn(M, K, device=dev)
        b = torch.randn(N, K, device=dev) / K ** 0.5
        ah, al = torch.empty(M, ldk, dtype=torch.bfloat16, device=dev), torch.empty(M, ldk, dtype=torch.bfloat16, device=dev)
        bh, bl = torch.empty(N, ldk, dtype=torch.bfloat16, device=dev), torch.empty(N, ldk, dtype=torch.bfloat16, device=dev)
        c = torch.empty(M, N, device=dev)
        check(L.pcrcg_split_bf16_dev(a.data_ptr(), K, M, K, ah.data_ptr(), al.data_ptr(), ldk, st))
        check(L.pcrcg_split_bf16_dev(b.data_ptr(), K, N, K, bh.data_ptr(), bl.data_ptr(), ldk, st))
        if "--stats" in sys.argv:        # epilogue InstanceNorm statistics, 16 segments
            seg = torch.linspace(0, M, 17, device=dev).to(torch.int32)
            acc = torch.zeros(16, 2, N, dtype=torch.float64, device=dev)
            run = lambda: check(L.pcrcg_gemm_bf16x3_stats_dev(ah.data_ptr(), al.data_ptr(), bh.data_ptr(), bl.data_ptr(), ldk, c.data_ptr(), N, M, N, K,
                                                              None, seg.data_ptr(), 16, acc.data_ptr(), st))
        else:
            run = lambda: check(L.pcrcg_gemm_bf16x3_dev(ah.data_ptr(), al.data_ptr(), bh.data_ptr(), bl.data_ptr(), ldk, c.data_ptr(), N, M, N, K, None, st))
        for _ in range(3):
            run()
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        fl = 2.0 * M * N * K
        byts = 4.0 * M * ldk + 4.0 * N * ldk + 4.0 * M * N
        ref = a[:256].double() @ b.double().t()
        err = float((c[:256].double() - ref).abs().max() / ref.abs().max())
        print(f"{name:24s} M={M:7d} N={N:5d} K={K:5d}  {ms:8.3f} ms  useful {fl / ms / 1e9:7.1f} TF/s  pipe {3 * fl / ms / 1e9:7.1f} TF/s  "
              f"{byts / ms / 1e6:7.0f} GB/s  hbm-floor {byts / 6548.5e6:6.3f} ms  err {err:.1e}", flush=True)


if __name__ == "__main__":
    main()
