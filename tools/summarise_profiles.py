#!/usr/bin/env python3
"""Turns the scratch ncu output of tools/profile_gpu.sh (gpurun_out/prof/) into the tracked summaries under profiles/.
usage: python tools/summarise_profiles.py r01"""
import collections
import csv
import os
import subprocess
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
SRC, DST = "gpurun_out/prof", "profiles"
os.makedirs(DST, exist_ok=True)

# 1. launch list (shares of the step)
rows = [r for r in csv.reader(open(f"{SRC}/{R}_launches_bench.csv")) if len(r) > 10 and r[0].isdigit()]
with open(f"{DST}/{R}_launches_bench.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(["id", "kernel", "grid", "block", "gpu__time_duration.sum_ns"])
    for r in rows:
        w.writerow([r[0], r[4].split("(")[0].replace("void ", ""), r[8], r[7], r[-1].replace(",", "")])
agg = collections.OrderedDict()
for r in rows:
    k = r[4].split("(")[0].replace("void ", "")
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[-1].replace(",", ""))
tot = sum(v for _, v in agg.values())
with open(f"{DST}/{R}_launch_shares.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --steps 3 --warmup 3 (cold-cache, serialised)\n")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:70s} n={n:3d} total={v / 1e6:9.3f} ms  share={100 * v / tot:5.1f}%\n")
print(open(f"{DST}/{R}_launch_shares.txt").read())

# 2. full capture of the dominant kernel: selected raw metrics
raw = subprocess.run(["ncu", "-i", f"{SRC}/{R}_k_sis_fused.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, vals = rr[0], rr[1], rr[2]
keep = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
with open(f"{DST}/{R}_k_sis_fused_ncu_full.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    for h, u, v in zip(hdr, units, vals):
        if h in keep:
            w.writerow([h, u, v])
            print(f"{h:90s} {u:16s} {v}")

# 3. row-path kernels
rows = [r for r in csv.reader(open(f"{SRC}/{R}_rows_kernels.csv")) if len(r) > 10 and r[0].isdigit()]
by = collections.OrderedDict()
for r in rows:
    by.setdefault((r[0], r[4].split("(")[0].replace("void ", ""), r[8]), {})[r[12]] = (r[14].replace(",", ""), r[13])
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
with open(f"{DST}/{R}_rows_kernels_ncu.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(["id", "kernel", "grid", "time_ms", "dram_read_GB", "dram_write_GB", "dram_GBps", "dram_pct", "fp64_pipe_pct", "issue_active_pct", "regs", "warps_active_pct", "warp_inst"])
    for (i, k, g), m in by.items():
        t, tu = float(m["gpu__time_duration.sum"][0]), m["gpu__time_duration.sum"][1]
        t_ms = t / 1e6 if tu.startswith("n") else (t / 1e3 if tu.startswith("u") else t)
        rd = float(m["dram__bytes_read.sum"][0]) * mult.get(m["dram__bytes_read.sum"][1], 1)
        wr = float(m["dram__bytes_write.sum"][0]) * mult.get(m["dram__bytes_write.sum"][1], 1)
        w.writerow([i, k, g, f"{t_ms:.4f}", f"{rd / 1e9:.4f}", f"{wr / 1e9:.4f}", f"{(rd + wr) / 1e9 / (t_ms / 1e3):.1f}",
                    m["dram__throughput.avg.pct_of_peak_sustained_elapsed"][0], m["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"][0],
                    m["smsp__issue_active.avg.pct_of_peak_sustained_active"][0], m["launch__registers_per_thread"][0],
                    m["sm__warps_active.avg.pct_of_peak_sustained_active"][0], m["smsp__inst_executed.sum"][0]])
print(open(f"{DST}/{R}_rows_kernels_ncu.csv").read())
