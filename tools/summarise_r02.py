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

# ---- launch list of the bench command (shares of the step) ------------------------------------------------------------
path = f"{SRC}/{T}_launches_bench.csv"
if os.path.exists(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    with open(f"{DST}/{T}_launches_bench.csv", "w") as f:
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
    with open(f"{DST}/{T}_launch_shares.txt", "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --steps 3 --warmup 3 (cold-cache, serialised; "
                "all configurations of the bench line: C2 steps, strong-scaling extra, C3, C4, C5, file emission)\n")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:78s} n={n:3d} total={v / 1e6:9.3f} ms  share={100 * v / tot:5.1f}%\n")
    print(open(f"{DST}/{T}_launch_shares.txt").read())

# ---- full captures: selected raw metrics + instruction budgets from the source pages --------------------------------------
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
CAPTURES = {"fused_c2": ("C2", 200_000_000, 1), "staged_c3": ("C3", 1 << LG, 32), "staged_c4": ("C4", 1 << LG, 64), "staged_c5": ("C5_estimators", 1 << (LG - 4), 1000), "rows_c5": ("C5_rows_kernel", 1 << (LG - 4), 1000)}
costs = {}
for name, (label, particles, steps) in CAPTURES.items():
    raw, src = f"{SRC}/{T}_{name}_raw.csv", f"{SRC}/{T}_{name}_source.csv"
    if not (os.path.exists(raw) and os.path.exists(src)):
        continue
    rr = list(csv.reader(open(raw)))
    hdr, units, vals = rr[0], rr[1], rr[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    with open(f"{DST}/{T}_{name}_ncu_full.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for k in KEEP:
            if k in m:
                w.writerow([k, m[k][1], m[k][0]])
    srows = list(csv.reader(open(src)))
    sh = srows[1]
    i_src, i_exec = sh.index("Source"), sh.index("Instructions Executed")
    tot = fp64 = 0
    for r in srows[2:]:
        if len(r) <= i_exec or not r[0].startswith("0x"):
            continue
        n = int(r[i_exec])
        toks = r[i_src].split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "")
        tot += n
        if op.split(".")[0] in ("DADD", "DFMA", "DMUL", "DSETP", "DMNMX"):
            fp64 += n
    dram = nbytes(m, "dram__bytes_read.sum") + nbytes(m, "dram__bytes_write.sum")
    costs[label] = {"kernel": m.get("Kernel Name", ("", ""))[0], "particles": particles, "steps_per_particle": steps,
                    "warp_instructions": tot, "fp64_warp_instructions": fp64,
                    "instructions_per_particle": tot * 32 / particles, "fp64_instructions_per_particle": fp64 * 32 / particles,
                    "issue_slots_per_particle": (tot + fp64) * 32 / particles, "issue_slots_per_particle_step": (tot + fp64) * 32 / particles / steps,
                    "dram_bytes_per_launch": dram, "dram_bytes_per_particle": dram / particles, "time_ms": ms(m),
                    "issue_active_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
                    "fp64_pipe_pct": float(m["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"][0]),
                    "registers": int(float(m["launch__registers_per_thread"][0])), "source": f"profiles/{T}_{name}_ncu_full.csv"}
    print(label, json.dumps(costs[label]))
# row path of C5 (emitting runs): DRAM bytes of all its kernels per particle, from the per-kernel pass ("C5 rows" of configs_once)
rows_bytes, in_c5 = 0.0, False
for (i, k, g, b), m in by.items():
    if "k_sis_rows<models::hmm_model>" in k and not in_c5 and rows_bytes == 0.0:
        in_c5 = True
    if in_c5:
        rows_bytes += nbytes(m, "dram__bytes_read.sum") + nbytes(m, "dram__bytes_write.sum")
        if "k_merge_columns" in k:
            break
if rows_bytes:
    costs["C5"] = {"dram_bytes_per_particle": rows_bytes / (1 << (LG - 4)), "particles": 1 << (LG - 4),
                   "source": f"profiles/{T}_config_kernels_ncu.csv (k_sis_rows ... k_merge_columns of the forced-rows hmm<1000> pass)"}
if costs:
    with open(f"{DST}/kernel_costs.json", "w") as f:
        json.dump({"tag": T, "how": "tools/profile_r02.sh + tools/summarise_r02.py: ncu --set full --clock-control none --import-source on, one launch each; "
                                    "instructions = executed warp instructions of the source page x 32 / particles; issue slots = instructions + "
                                    "FP64-pipe instructions (an FP64 instruction holds the issue port for two cycles)", **costs}, f, indent=1)
    print("wrote profiles/kernel_costs.json")
