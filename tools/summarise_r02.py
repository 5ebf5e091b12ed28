#!/usr/bin/env python3
"""Turns the scratch ncu output of tools/profile_r02.sh (gpurun_out/prof2/) into tracked summaries under profiles/.
usage: python tools/summarise_r02.py <tag> [particles log2 of configs_once, default 24]"""
import collections
import csv
import json
import os
import sys

T = sys.argv[1]
LG = int(sys.argv[2]) if len(sys.argv) > 2 else 24
SRC, DST = "gpurun_out/prof2", "profiles"
MULT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def num(v):
    return float(v.replace(",", ""))


def kernels(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    by = collections.OrderedDict()
    for r in rows:
        by.setdefault((int(r[0]), r[4].split("(")[0].replace("void ", ""), r[8], r[7]), {})[r[12]] = (r[14], r[13])
    return by


def ms(m):
    v, u = num(m["gpu__time_duration.sum"][0]), m["gpu__time_duration.sum"][1]
    return v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else (v if u.startswith("m") else v * 1e3))


def nbytes(m, k):
    return num(m[k][0]) * MULT.get(m[k][1], 1)


by = kernels(f"{SRC}/{T}_config_kernels.csv")
out = f"{DST}/{T}_config_kernels_ncu.csv"
with open(out, "w") as f:
    w = csv.writer(f)
    w.writerow(["id", "kernel", "grid", "block", "time_ms", "dram_read_GB", "dram_write_GB", "dram_GBps", "dram_pct", "fp64_pipe_pct",
                "issue_active_pct", "regs", "warps_active_pct", "warp_inst", "thread_inst", "eligible_warps", "smem_bank_conflicts",
                "smem_ld_inst", "smem_st_inst"])
    for (i, k, g, b), m in by.items():
        t = ms(m)
        rd, wr = nbytes(m, "dram__bytes_read.sum"), nbytes(m, "dram__bytes_write.sum")
        g_ = lambda key: m.get(key, ("", ""))[0].replace(",", "")
        w.writerow([i, k, g, b, f"{t:.4f}", f"{rd / 1e9:.4f}", f"{wr / 1e9:.4f}", f"{(rd + wr) / 1e9 / (t / 1e3):.1f}",
                    g_("dram__throughput.avg.pct_of_peak_sustained_elapsed"), g_("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                    g_("smsp__issue_active.avg.pct_of_peak_sustained_active"), g_("launch__registers_per_thread"),
                    g_("sm__warps_active.avg.pct_of_peak_sustained_active"), g_("smsp__inst_executed.sum"), g_("smsp__thread_inst_executed.sum"),
                    g_("smsp__warps_eligible.avg.per_cycle_active"), g_("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
                    g_("smsp__inst_executed_op_shared_ld.sum"), g_("smsp__inst_executed_op_shared_st.sum")])
print(open(out).read())
