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
