#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv          > profiles/rNN_launches.txt
  python profiles/summarize.py kernel   gpurun_out/prof_dp.ncu-rep       > profiles/rNN_dp_full.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "inst_executed",
        "thread_inst_executed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sass__inst_executed_register_spilling",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sass__inst_executed_global_loads",
        "sass__inst_executed_global_stores", "sass__inst_executed_shared_loads", "smsp__inst_executed_op_tma_ld.sum",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio"]
STALLS = "smsp__average_warps_issue_stalled_"


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        agg.setdefault(row["Kernel Name"], []).append(float(row["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    for k, v in agg.items():
        print("%-80s launches=%3d  total=%10.3f ms  share=%5.1f%%" % (k[:80], len(v), sum(v) / 1e6, 100 * sum(v) / tot))


def kernel(path):
    raw = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"]).decode().splitlines()
    rows = list(csv.reader(raw))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("=" * 100)
        print(r[idx["Kernel Name"]], " grid", r[idx["Grid Size"]], " block", r[idx["Block Size"]])
        for k in KEYS:
            if k in idx:
                print("  %-72s %16s %s" % (k, r[idx[k]], units[idx[k]]))
        for h in hdr:
            if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio"):
                try:
                    v = float(r[idx[h]].replace(",", ""))
                except ValueError:
                    continue
                if v >= 0.05:
                    print("  stall %-66s %16.3f" % (h[len(STALLS):-len("_per_issue_active.ratio")], v))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
