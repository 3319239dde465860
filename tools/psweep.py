#!/usr/bin/env python
"""Pairs per step sweep (P in {1, 4, 16, 32, 64}) of the 3DMatch-shaped workload on one GPU: device-resident
throughput of the synchronous call (run_device: the latency of ONE request of P pairs) and of the pipelined submission
(submit_device, two steps in flight).  Writes a markdown table.  Usage: python tools/psweep.py [out.md]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from pcrcg_b200 import pipeline  # noqa: E402
from pcrcg_b200._lib import lib  # noqa: E402

dev = torch.device("cuda:0")
cfg, limits = bench.workload_config("3dmatch")
all_pairs = bench.make_pairs("3dmatch", 64, 0)
path = pipeline.FeaturePath(cfg, limits, device=dev, seed=0)
L = lib()
rows = []
for P in (1, 4, 16, 32, 64):
    pts_np, lens_np = pipeline.stack_pairs(all_pairs[:P])
    pts, lens = torch.from_numpy(pts_np).to(dev), torch.from_numpy(lens_np).to(dev)
    K = max(6, min(40, 256 // P))
    for _ in range(3):
        path.run_device(pts, lens)
    hs = [path.submit_device(pts, lens) for _ in range(2)]
    hs[-1].result()
    torch.cuda.synchronize()
    n0 = L.pcrcg_launch_count()
    t0 = time.perf_counter()
    for _ in range(K):
        path.run_device(pts, lens)
    torch.cuda.synchronize()
    t_sync = (time.perf_counter() - t0) / K
    launches = (L.pcrcg_launch_count() - n0) // K
    t0 = time.perf_counter()
    hs = []
    for i in range(K):
        if i >= 2:
            hs[i - 2].ready.synchronize()
            hs[i - 2] = None
        hs.append(path.submit_device(pts, lens))
    hs[-1].result()
    torch.cuda.synchronize()
    t_pipe = (time.perf_counter() - t0) / K
    rows.append((P, pts.shape[0], launches, 1000 * t_sync, P / t_sync, 1000 * t_sync / P, 1000 * t_pipe, P / t_pipe))
    print(rows[-1], flush=True)

out = ["| pairs per step | points (level 0) | kernel launches / step | run_device ms / step | pairs/s | ms / pair | pipelined ms / step | pipelined pairs/s |",
       "|---:|---:|---:|---:|---:|---:|---:|---:|"]
for r in rows:
    out.append(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]:.2f} | {r[4]:.0f} | {r[5]:.2f} | {r[6]:.2f} | {r[7]:.0f} |")
text = "\n".join(out) + "\n"
print(text)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
