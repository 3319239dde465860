#!/usr/bin/env python
"""Digest of an `ncu --page raw --csv` export: one line per profiled launch with the metrics the
roofline discussion needs.  Usage: ncu_digest.py raw.csv [out.md]"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur_us", 1e-3),
    ("dram__bytes_read.sum", "dram_rd_MB", 1e-6),
    ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%", 1),
    ("lts__t_bytes.sum", "l2_MB", 1e-6),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%", 1),
    ("smsp__inst_executed.sum", "inst_M", 1e-6),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__grid_size", "grid", 1),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "st_long%", 1),
    ("smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "st_short%", 1),
    ("smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "st_math%", 1),
    ("smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "st_lg%", 1),
    ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "st_bar%", 1),
]


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = ["| kernel | " + " | ".join(k[1] for k in KEYS) + " |", "|---|" + "---:|" * len(KEYS)]
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[col["Kernel Name"]].split("(")[0].replace("pcrcg::", "").replace("void ", "")[:60]
        vals = []
        for key, _, scale in KEYS:
            if key not in col:
                vals.append("-")
                continue
            v = num(r[col[key]])
            u = units[col[key]]
            if v is None:
                vals.append("-")
                continue
            if key == "gpu__time_duration.sum":
                v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
            elif "bytes" in key:
                v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
            else:
                v = v * scale
            vals.append(f"{v:.1f}" if abs(v) < 1e5 else f"{v:.3g}")
        out.append(f"| {name} | " + " | ".join(vals) + " |")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
