#!/usr/bin/env python
"""profiles/traffic.json from an ncu launch list of ONE full hot-path step with DRAM counters:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -s <launches before the step> -c <launches of one step> --csv --log-file launches.csv python bench.py ...

Per kernel class of bench.py (same names as its `kernels` table): DRAM bytes per step, launches per step, time under ncu.
k_gemm_bf16x3 launches are split by their position: the launch that follows a k_kpconv_aggregate* kernel is that KPConv's
weight contraction ("kpconv_contraction"), every other one a unary Linear ("linear").
Usage: traffic_from_ncu.py launches.csv out.json "<source description>" [first_id last_id]"""
import csv
import json
import re
import sys
from collections import OrderedDict

CLASSES = [
    ("radius", r"k_cell_|k_radius_|k_rbbox|k_grid_meta"),
    ("subsample", r"k_bbox|k_keys|k_insert|k_rs_|k_heads|k_bary|k_order|k_origin|k_outlens"),
    ("kpconv_fused", r"k_kpconv_fused"),
    ("kpconv_aggregate", r"k_kpconv_aggregate|k_row_positive|k_split_rows|k_split_w_tail1"),
    ("norm_act", r"k_colstats|k_norm_act|k_bias_act"),
    ("pool", r"k_max_pool|k_closest_pool"),
    ("projection", r"k_project"),
]


def main():
    lines = [l for l in open(sys.argv[1], newline="") if l.startswith('"')]
    launches = OrderedDict()
    for r in csv.DictReader(lines):
        i = int(r["ID"])
        e = launches.setdefault(i, {"name": re.sub(r"^void\s+", "", re.sub(r"\(.*", "", r["Kernel Name"])).replace("pcrcg::", "")})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        if r["Metric Name"].startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            e["bytes"] = e.get("bytes", 0.0) + v
        elif r["Metric Name"] == "gpu__time_duration.sum":
            e["ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    ids = list(launches)
    if len(sys.argv) > 5:
        ids = [i for i in ids if int(sys.argv[4]) <= i <= int(sys.argv[5])]
    per = {}
    prev_agg = False
    for i in ids:
        e = launches[i]
        name = e["name"]
        if re.search(r"k_gemm_bf16x3|k_sgemm", name):
            cls = "kpconv_contraction" if prev_agg else "linear"
            prev_agg = False
        elif re.search(r"k_split_bf16", name):
            cls = "linear"          # operand splits of either kind: small, counted with the Linears; do not reset prev_agg
        else:
            cls = next((c for c, pat in CLASSES if re.search(pat, name)), None)
            prev_agg = bool(re.search(r"k_kpconv_aggregate", name)) or (prev_agg and cls in (None,))
        if cls is None:
            continue
        p = per.setdefault(cls, {"launches": 0, "bytes": 0.0, "ms": 0.0})
        p["launches"] += 1
        p["bytes"] += e.get("bytes", 0.0)
        p["ms"] += e.get("ms", 0.0)
    out = {"source": sys.argv[3] if len(sys.argv) > 3 else sys.argv[1], "unit": "DRAM bytes per step (dram__bytes_read.sum + dram__bytes_write.sum)",
           "per_class": {k: {"launches_per_step": v["launches"], "dram_bytes_per_step": v["bytes"], "ncu_ms_per_step": round(v["ms"], 4)}
                         for k, v in per.items()}}
    for k, v in per.items():
        out[k] = v["bytes"]
    if all(k in per for k in ("kpconv_aggregate", "kpconv_contraction")):
        out["kpconv"] = per["kpconv_aggregate"]["bytes"] + per["kpconv_contraction"]["bytes"] + per.get("kpconv_fused", {"bytes": 0.0})["bytes"]
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(json.dumps(out["per_class"], indent=1))


if __name__ == "__main__":
    main()
