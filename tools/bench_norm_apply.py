#!/usr/bin/env python
"""Micro-benchmark of the InstanceNorm APPLY kernel alone (statistics given), per launch-shape variant
(pcrcg_set_option "norm_variant") on the hot path's tensor shapes (16 stacked pairs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pcrcg_b200 import ops  # noqa: E402
from pcrcg_b200._lib import lib  # noqa: E402

L = lib()
dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
CASES = [  # n, c, shortcut, planes_only
    (646588, 64, False, True), (646588, 64, False, False), (646588, 128, False, False), (646588, 256, True, False),
    (156404, 128, False, True), (156404, 512, True, False), (42419, 1024, True, False)]
for n, c, has_sc, po in CASES:
    x = torch.randn(n, c, device=dev)
    sc = torch.randn(n, c, device=dev) if has_sc else None
    seg = torch.linspace(0, n, 17, device=dev).to(torch.int32)
    st = ops.column_stats(x, seg)
    x._pcrcg_stats = (st[0], st[1], seg, 1e-5)
    if has_sc:
        s2 = ops.column_stats(sc, seg)
        sc._pcrcg_stats = (s2[0], s2[1], seg, 1e-5)
    for variant in range(5):
        L.pcrcg_set_option(b"norm_variant", variant)
        run = lambda: ops.instance_norm_act(x, seg, 0.1, shortcut=sc, shortcut_norm=has_sc, emit_split=True, emit_rowpos=po, planes_only=po)
        for _ in range(2):
            run()
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[2]
        byts = n * c * 4 * (1 + (1 if has_sc else 0) + (0 if po else 1) + 1)
        print(f"[{n}x{c}] sc={has_sc} planes_only={po} variant={variant}: {ms:.3f} ms  {byts / ms / 1e6:.0f} GB/s", flush=True)
