"""CPU: bench.py's HeadlineGuard -- the headline JSON line reaches stdout when the tail of the run (other workloads, parity check)
hangs, exceeds its deadline or the process is terminated, and is printed exactly once otherwise."""
import json
import os
import signal
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, time, json
sys.path.insert(0, {root!r})
import bench
g = bench.HeadlineGuard()
g.arm({{"metric": "m", "value": 1.5}}, {deadline})
print("armed", file=sys.stderr, flush=True)
mode = {mode!r}
if mode == "finish":
    time.sleep(0.2)
    assert g.finish({{"metric": "m", "value": 1.5, "other_workloads": []}})
    time.sleep(0.3)
else:
    time.sleep(30)          # a tail that never finishes (the guard ends the process)
    print("unreachable")
"""


def _run(mode, deadline, send_term=False):
    p = subprocess.Popen([sys.executable, "-c", SCRIPT.format(root=ROOT, deadline=deadline, mode=mode)], stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True)
    if send_term:
        assert p.stderr.readline().strip() == "armed"
        time.sleep(0.3)
        p.send_signal(signal.SIGTERM)
    out, _ = p.communicate(timeout=20)
    return p.returncode, [json.loads(l) for l in out.splitlines() if l.strip()]


def test_complete_line_printed_once():
    rc, lines = _run("finish", 5.0)
    assert rc == 0 and len(lines) == 1 and lines[0]["other_workloads"] == [] and "incomplete" not in lines[0]


def test_deadline_prints_headline_and_ends_the_process():
    t0 = time.time()
    rc, lines = _run("hang", 0.5)
    assert rc == 0 and time.time() - t0 < 15
    assert len(lines) == 1 and lines[0]["value"] == 1.5 and "did not finish" in lines[0]["incomplete"]


def test_sigterm_prints_headline():
    rc, lines = _run("hang", 60.0, send_term=True)
    assert rc == 143
    assert len(lines) == 1 and lines[0]["value"] == 1.5 and "signal 15" in lines[0]["incomplete"]


def test_class_summary_reproduces_the_committed_bench_line():
    """bench.algorithmic_work + bench.summarise_classes (pure host code) fed with the per-class times of the committed round-2
    line give that line's GB/s, TFLOP/s and roofline entries back"""
    import torch
    import bench
    from pcrcg_b200 import blocks
    ref = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_n1.json")))
    cfg = blocks.indoor_config()
    N, lim = ref["config"]["points_per_level"], ref["config"]["limits"]
    enc = type("E", (), {})()
    enc.encoder_blocks = blocks.KPEncoder(cfg).encoder_blocks
    empty = torch.empty(0, 1, dtype=torch.int32)
    batch = dict(points=[torch.empty(n, 3) for n in N], stack_lengths=[torch.empty(64, dtype=torch.int32)],
                 neighbors=[torch.empty(n, w, dtype=torch.int32) for n, w in zip(N, lim)],
                 pools=[torch.empty(N[i + 1], lim[i], dtype=torch.int32) for i in range(3)] + [empty],
                 upsamples=[torch.empty(N[i], lim[i], dtype=torch.int32) for i in range(3)] + [empty])
    work, levels = bench.algorithmic_work(batch, cfg, lim, enc)
    assert levels == N
    K = ref["steps"]
    names = {"subsample": "subsample", "radius": "radius_query", "kpconv_aggregate": "kpconv_aggregate", "kpconv_fused": "kpconv_fused",
             "kpconv_contraction": "gemm", "linear": "linear", "norm_act": "norm_act", "pool": "pool"}
    prof = {names[k]: (v["ms_per_step"] * K, v["scopes_per_step"] * K) for k, v in ref["kernels"].items() if k in names}
    kernels, roof = bench.summarise_classes(prof, work, K)
    for k in names:
        for key in ("GB/s", "TFLOP/s"):
            if key in ref["kernels"][k]:
                assert abs(kernels[k][key] - ref["kernels"][k][key]) <= 2e-3 * ref["kernels"][k][key], (k, key)
    assert roof["kernel"] == ref["roofline"]["kernel"] == "kpconv_aggregate" and roof["bound"] == "hbm"
    assert roof["algorithmic_bytes"] == ref["roofline"]["algorithmic_bytes"]
    assert abs(roof["frac"] - ref["roofline"]["frac"]) < 1e-3
    # the aggregation is an L2 gather: ~25 GB of neighbour rows per step for the two-kernel layers, ~14 GB for the fused ones
    assert 4000 < kernels["kpconv_aggregate"]["gathered_GB/s"] < 5000 and 3000 < kernels["kpconv_fused"]["gathered_GB/s"] < 4500
    assert roof["gathered_GB/s"] == kernels["kpconv_aggregate"]["gathered_GB/s"]
