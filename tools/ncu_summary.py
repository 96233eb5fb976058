#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small, committed summaries under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches.txt
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/r1_dock_step_f32.txt [dtype-key]

`full` also updates profiles/roofline_traffic.json (dram bytes per launch of the step kernel; bench.py reads it).
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__grid_size", "launch__block_size", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "sm__cycles_elapsed.max",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
    "smsp__inst_executed_pipe_uniform.sum", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
]


def launches(src, dst):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# source: %s ; total %.1f us over %d launches\n" % (src, tot / 1e3, sum(len(v) for v in agg.values())))
        f.write("%-90s %6s %12s %10s %7s\n" % ("kernel", "n", "sum_us", "mean_us", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("%-90s %6d %12.1f %10.2f %7.3f\n" % (k[:90], len(v), sum(v) / 1e3, sum(v) / len(v) / 1e3, sum(v) / tot))
    print(open(dst).read())


def full(src, dst, key=None):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on ; source: %s\n" % src)
        f.write("# kernels: %s\n" % sorted(set(r[name_i] for r in data)))
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                f.write("%-78s %-14s %s\n" % (m, units[i], " ".join(r[i] for r in data)))
        f.write("# warp stall reasons (average warps stalled per issue-active cycle), values > 0.1\n")
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                v = [float(r[i]) for r in data]
                if max(v) > 0.1:
                    f.write("  %-40s %s\n" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], " ".join("%.2f" % x for x in v)))
    print(open(dst).read())
    if key:
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tr = [float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]] for r in data]
        path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        d = json.load(open(path)) if os.path.exists(path) else {}
        d[key] = sum(tr) / len(tr)
        d[key + "_source"] = os.path.basename(dst)
        json.dump(d, open(path, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
