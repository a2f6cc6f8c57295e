#!/usr/bin/env python
"""Condense an ncu report into one CSV row per kernel launch (the metrics the profiles/README.md tables quote).

    python profiles/extract.py gpurun_out/x.ncu-rep profiles/r1b_ncu_full_xxx.csv
"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    stalls = [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
    cols = [w for w in WANT if w in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols + ["top_stalls(warps per issue slot)"])
        w.writerow([units[hdr.index(c)] for c in cols] + [""])
        for r in rows[2:]:
            st = sorted(((float(r[hdr.index(h)].replace(",", "") or 0), h.split("issue_stalled_")[1].split("_per")[0]) for h in stalls), reverse=True)[:4]
            w.writerow([r[hdr.index(c)] for c in cols] + ["; ".join("%s %.2f" % (n, v) for v, n in st)])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
