#!/usr/bin/env python
"""One hot-path step between cudaProfilerStart/Stop, for `ncu --profile-from-start off`.
Usage: ncu ... python tools/prof_step.py [--pairs P] [--workload 3dmatch]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from pcrcg_b200 import pipeline  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=4)
ap.add_argument("--workload", default="3dmatch")
a = ap.parse_args()
cfg, limits = bench.workload_config(a.workload)
pairs = bench.make_pairs(a.workload, a.pairs, 0)
pts, lens = pipeline.stack_pairs(pairs)
path = pipeline.FeaturePath(cfg, limits, device="cuda:0")
p, l = torch.from_numpy(pts).cuda(), torch.from_numpy(lens).cuda()
for _ in range(3):
    path.run_device(p, l)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
path.run_device(p, l)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step:", a.pairs, "pairs")
