#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU): one line per profiled launch with the metrics the roofline discussion uses.
usage: ncu_summary.py REPORT.ncu-rep [--json OUT]"""
import csv, io, json, subprocess, sys

KEYS = {
    "gpu__time_duration.sum": "us",
    "dram__bytes_read.sum": "dram_rd_MB",
    "dram__bytes_write.sum": "dram_wr_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "smsp__inst_executed.sum": "warp_inst",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "st_long_sb",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "st_barrier",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "st_short_sb",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "st_math",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "st_lg",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "st_mio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "st_wait",
}

def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        e = {"kernel": d.get("Kernel Name", "?").split("(")[0]}
        for k, short in KEYS.items():
            if k in d:
                try:
                    v = float(d[k].replace(",", ""))
                except ValueError:
                    continue
                u = units[hdr.index(k)]
                if short == "us":
                    v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                if short.endswith("_MB"):
                    v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                e[short] = round(v, 3)
        res.append(e)
    cols = ["kernel"] + list(KEYS.values())
    print(" | ".join(cols))
    for e in res:
        print(" | ".join(str(e.get(c, "")) for c in cols))
    if "--json" in sys.argv:
        json.dump(res, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)

main()
