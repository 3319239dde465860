#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full --page raw --csv` export of ONE hot-path step (tools/prof_step.py):
per kernel class (the classes bench.py times) the DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum), the number
of launches, bytes per launch of the class's dominant kernel, time under ncu and time-weighted tensor-pipe activity.
Usage: traffic_from_ncu.py raw.csv out.json "<source description>" """
import csv
import json
import re
import sys

CLASSES = [
    ("radius", r"k_cell_|k_radius_query|k_rbbox|k_knn|k_point2node"),
    ("subsample", r"k_bbox|k_keys|k_insert|k_rs_|k_heads|k_bary|k_order|k_scan|k_cloud_starts|k_compact|k_mark"),
    ("kpconv_aggregate", r"k_kpconv_aggregate|k_row_positive"),
    ("gemm", r"k_gemm_bf16x3|k_sgemm|k_split_bf16"),
    ("norm_act", r"k_colstats|k_norm_act|k_bias_act|k_softmax|k_l2norm|k_descriptor_head"),
    ("pool", r"k_max_pool|k_closest_pool|k_edge_max"),
]


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return 0.0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    u = lambda name: units[col[name]]
    scale_b = lambda name: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u(name), 1.0)
    scale_t = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u("gpu__time_duration.sum"), 1e-6)
    per = {}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[col["Kernel Name"]]
        cls = next((c for c, pat in CLASSES if re.search(pat, name)), None)
        if cls is None:
            continue
        byts = num(r[col["dram__bytes_read.sum"]]) * scale_b("dram__bytes_read.sum") + \
            num(r[col["dram__bytes_write.sum"]]) * scale_b("dram__bytes_write.sum")
        ms = num(r[col["gpu__time_duration.sum"]]) * scale_t
        tp = num(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]) if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in col else 0.0
        short = re.sub(r"^void\s+", "", re.sub(r"[<(].*", "", name)).replace("pcrcg::", "")
        e = per.setdefault(cls, {"launches": 0, "bytes": 0.0, "ms": 0.0, "tensor_ms": 0.0, "by_kernel": {}})
        e["launches"] += 1
        e["bytes"] += byts
        e["ms"] += ms
        e["tensor_ms"] += tp * ms
        k = e["by_kernel"].setdefault(short, [0, 0.0, 0.0])
        k[0] += 1; k[1] += byts; k[2] += ms
    out = {"source": sys.argv[3] if len(sys.argv) > 3 else sys.argv[1], "per_class": {}}
    for cls, e in per.items():
        dom = max(e["by_kernel"].items(), key=lambda kv: kv[1][2])
        out["per_class"][cls] = {
            "launches_profiled_per_step": e["launches"], "dram_bytes_per_step": e["bytes"], "ncu_time_ms_per_step": e["ms"],
            "dram_GBps_under_ncu": e["bytes"] / e["ms"] / 1e6 if e["ms"] else 0.0,
            "tensor_pipe_active_pct_timeweighted": e["tensor_ms"] / e["ms"] if e["ms"] else 0.0,
            "dominant_kernel": dom[0], "dominant_kernel_launches": dom[1][0], "dominant_kernel_dram_bytes_per_launch": dom[1][1] / dom[1][0]}
        out[cls] = dom[1][1] / dom[1][0]          # what bench.py reports as roofline.traffic (per launch of the dominant kernel)
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k not in ("per_class", "source")}, indent=1))


if __name__ == "__main__":
    main()
