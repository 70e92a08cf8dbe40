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


def traffic(path, bench_json, out_json):
    """per-kernel DRAM bytes of one bench step (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,
    gpu__time_duration.sum) set against the algorithmic bytes the same run's bench line reports"""
    import json
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    ids = collections.defaultdict(set)
    for row in csv.DictReader(lines):
        d = agg.setdefault(row["Kernel Name"], collections.defaultdict(float))
        ids[row["Kernel Name"]].add(row["ID"])
        v = float(row["Metric Value"].replace(",", ""))
        unit = row.get("Metric Unit", "")
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0,
                "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1.0)
        d[row["Metric Name"]] += v * mult
    # bench.py runs the batch several times (device-timed pass, end-to-end pass, ...): report ONE pass
    passes = max(len(v) for v in ids.values())
    for d in agg.values():
        for k in d:
            d[k] /= passes
    b = json.load(open(bench_json))
    sets = {"indel": [0.0, 0.0], "jump": [0.0, 0.0]}
    print("# ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none; average of the %d passes bench.py makes over the batch" % passes)
    for k, d in agg.items():
        rd, wr = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
        print("%-70s %9.3f ms  read %8.3f GB  write %8.3f GB" % (k[:70], d["gpu__time_duration.sum"] / 1e6, rd / 1e9, wr / 1e9))
        if "dpTagKernel<2" in k or "dpPairKernel<2" in k or "dpKernel<2" in k:
            sets["indel"][0] += rd; sets["indel"][1] += wr
        if "dpTagKernel<1" in k or "dpPairKernel<1" in k or "dpKernel<1" in k:
            sets["jump"][0] += rd; sets["jump"][1] += wr
    algo = float(b["roofline"]["algorithmic_bytes_per_launch_set"])
    tot = sum(sets["indel"])
    print("large-indel set: read %.3f GB + write %.3f GB = %.3f GB ; algorithmic bytes %.3f GB ; traffic/algorithmic = %.4f"
          % (sets["indel"][0] / 1e9, sets["indel"][1] / 1e9, tot / 1e9, algo / 1e9, tot / algo))
    jc = float(b["stages"]["jump_cells"])
    print("jump set:        read %.3f GB + write %.3f GB = %.3f GB ; cells (1 B each) %.3f G ; traffic/cells = %.4f"
          % (sets["jump"][0] / 1e9, sets["jump"][1] / 1e9, sum(sets["jump"]) / 1e9, jc / 1e9, sum(sets["jump"]) / jc))
    json.dump({"source": "ncu dram__bytes_{read,write}.sum over the dpTagKernel<KIND_INDEL,RT> launches of one step, "
                         "bench.py --loci %d (%s)" % (b["config"]["loci_per_gpu"], out_json.replace(".json", ".txt")),
               "dram_bytes_read_plus_write": tot, "algorithmic_bytes": algo, "traffic_per_algorithmic_byte": tot / algo,
               "jump_traffic_per_cell": sum(sets["jump"]) / jc,
               "note": "traffic above the algorithmic bytes is strip padding (rows up to the next multiple of 32*RT, the "
                       "31-step skew) and the traceback's reads of the pointer stream"}, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
