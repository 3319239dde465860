#!/usr/bin/env python
"""Device-resident loop variants of FeaturePath (one GPU): synchronous run_device vs pipelined submit_device with N steps in
flight, with / without the L2 flush.  Usage: python tools/pipeline_probe.py [pairs] [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from pcrcg_b200 import pipeline  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
cfg, limits = bench.workload_config("3dmatch")
pts_np, lens_np = pipeline.stack_pairs(bench.make_pairs("3dmatch", P, 0))
pts, lens = torch.from_numpy(pts_np).to(dev), torch.from_numpy(lens_np).to(dev)
path = pipeline.FeaturePath(cfg, limits, device=dev, seed=0)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for _ in range(3):
    path.run_device(pts, lens)
hs = [path.submit_device(pts, lens) for _ in range(3)]
hs[-1].result()
torch.cuda.synchronize()


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K, (time.perf_counter() - t0) * 1000 / K


def sync_loop(do_flush):
    def f():
        for _ in range(K):
            if do_flush:
                flush.fill_(0.0)
            path.run_device(pts, lens)
    return f


STAMPS = []


def piped(depth, do_flush):
    def f():
        hs = []
        del STAMPS[:]
        for i in range(K):
            if i >= depth:
                hs[i - depth].ready.synchronize()
                hs[i - depth] = None
            STAMPS.append(time.perf_counter())
            if do_flush:
                flush.fill_(0.0)
            hs.append(path.submit_device(pts, lens))
        hs[-1].result()
    return f


if os.environ.get("PROBE_SAMPLER"):
    smp = bench.ClockSampler(0)
    if os.environ["PROBE_SAMPLER"] == "smi":
        import pynvml as _p
        _p.nvmlInit = lambda: (_ for _ in ()).throw(RuntimeError("forced nvidia-smi"))
    smp.start()
    time.sleep(0.5)
    print("sampler mode:", smp.mode, flush=True)

for name, fn in (("sync, flush", sync_loop(True)), ("sync, no flush", sync_loop(False)), ("piped depth 1, flush", piped(1, True)),
                 ("piped depth 2, flush", piped(2, True)), ("piped depth 2, no flush", piped(2, False)), ("piped depth 3, flush", piped(3, True)), ("piped depth 2, flush, after sleep", piped(2, True))):
    for rep in range(2):
        st0 = torch.cuda.memory_stats()
        if "sleep" in name:
            time.sleep(0.3)
        ev, wall = timed(fn)
        if STAMPS and "sleep" in name:
            print("   host ms between submissions:", [round(1000 * (b - a), 1) for a, b in zip(STAMPS, STAMPS[1:])])
        st1 = torch.cuda.memory_stats()
        print(f"{name:28s} {ev:7.2f} ms/step (events)  {wall:7.2f} ms/step (host clock)  {1000 * P / ev:7.1f} pairs/s   "
              f"cudaMalloc +{st1['num_device_alloc'] - st0['num_device_alloc']} cudaFree +{st1['num_device_free'] - st0['num_device_free']} "
              f"reserved {st1['reserved_bytes.all.current'] / 2**30:.1f} GiB", flush=True)
