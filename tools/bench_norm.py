#!/usr/bin/env python
"""A/B micro-benchmark of the InstanceNorm kernels (column stats + apply) on hot-path tensor shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pcrcg_b200 import ops  # noqa: E402
from pcrcg_b200._lib import lib  # noqa: E402

L = lib()
dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for n, c in ((646588, 64), (646588, 256), (156404, 512), (42419, 1024)):
    x = torch.randn(n, c, device=dev)
    sc = torch.randn(n, c, device=dev)
    seg = torch.linspace(0, n, 17, device=dev).to(torch.int32)
    for v4 in (0, 1):
        L.pcrcg_set_option(b"norm_vectorised", v4)
        for split in (False, True):
            for _ in range(2):
                ops.instance_norm_act(x, seg, 0.1, shortcut=sc, shortcut_norm=True, emit_split=split)
            ts = []
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); ops.instance_norm_act(x, seg, 0.1, shortcut=sc, shortcut_norm=True, emit_split=split); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[2]
            byts = n * c * 4 * (2 + 2 + 1 + (1 if split else 0))     # 2 stats reads, 2 apply reads, 1 write (+ split planes)
            print(f"[{n}x{c}] v4={v4} split={split}: {ms:.3f} ms  {byts / ms / 1e6:.0f} GB/s (stats x2 + apply)", flush=True)
