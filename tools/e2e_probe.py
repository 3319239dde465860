#!/usr/bin/env python
"""Where does the end-to-end (host buffers) step lose time against the device-resident step?  Times 40 steps of:
device-resident loop; + pinned H2D of the inputs; + D2H of the result (the public submit_host path); D2H without the
allocator hold (result copied from a persistent device buffer)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from pcrcg_b200 import pipeline  # noqa: E402

cfg, limits = bench.workload_config("3dmatch")
pairs = bench.make_pairs("3dmatch", 32, 0)
pts, lens = pipeline.stack_pairs(pairs)
ph, lh = torch.from_numpy(pts).pin_memory(), torch.from_numpy(lens).pin_memory()
pd, ld = ph.cuda(), lh.cuda()
path = pipeline.FeaturePath(cfg, limits, device="cuda:0")
y, _ = path.run_device(pd, ld)
bufs = [torch.empty((y.shape[0] + 1024, y.shape[1]), dtype=torch.float32).pin_memory() for _ in range(2)]
K = 40


def timed(name, fn):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        fn(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name:46s} {32 * K / dt:8.1f} pairs/s  {1000 * dt / K:6.2f} ms/step", flush=True)


timed("device-resident", lambda i: path.run_device(pd, ld))
timed("+ H2D of the inputs", lambda i: path.run_device(ph.cuda(non_blocking=True), lh.cuda(non_blocking=True)))
pend = [None, None]


def full(i):
    if pend[i & 1] is not None:
        pend[i & 1].result()
    pend[i & 1] = path.submit_host(ph, lh, bufs[i & 1])


timed("submit_host (H2D + compute + pipelined D2H)", full)
copy_stream = torch.cuda.Stream()
keep = torch.empty_like(y)


def d2h_from_persistent(i):
    yy, _ = path.run_device(ph.cuda(non_blocking=True), lh.cuda(non_blocking=True))
    keep.copy_(yy)                                    # device copy on the compute stream; the result tensor is freed at once
    ev = torch.cuda.Event(); ev.record()
    with torch.cuda.stream(copy_stream):
        copy_stream.wait_event(ev)
        bufs[i & 1][:keep.shape[0]].copy_(keep, non_blocking=True)


timed("H2D + compute + D2H from a persistent buffer", d2h_from_persistent)
