#!/usr/bin/env python3
"""Hot regions of an `ncu --page source --csv` dump: executed warp instructions and stall samples per SASS line, grouped
into contiguous address ranges.  usage: python tools/ncu_source_hot.py <source.csv> [particles] [steps per particle]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) > iexec and r[ia].startswith("0x")]
base = int(body[0][ia], 16)
tot_exec = sum(int(r[iexec]) for r in body)
tot_samp = sum(int(r[isamp]) for r in body)
n_part = float(sys.argv[2]) if len(sys.argv) > 2 else None
steps = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
print(f"total warp instructions {tot_exec:.4g}, samples {tot_samp}")
if n_part:
    print(f"  = {tot_exec * 32 / n_part:.1f} thread-level instructions per particle, {tot_exec * 32 / n_part / steps:.2f} per particle-step")
# buckets of 64 instructions
B = 48
for b in range(0, len(body), B):
    chunk = body[b:b + B]
    ex = sum(int(r[iexec]) for r in chunk)
    sm = sum(int(r[isamp]) for r in chunk)
    if ex < 0.01 * tot_exec and sm < 0.01 * tot_samp:
        continue
    st = {}
    for i, h in stall_cols:
        st[h] = sum(int(r[i]) for r in chunk)
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    a0 = int(chunk[0][ia], 16) - base
    ops = {}
    for r in chunk:
        op = r[isrc].split()[0] if not r[isrc].split()[0].startswith("@") else r[isrc].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[iexec])
    topops = ", ".join(f"{k} {100 * v / max(ex, 1):.0f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:5])
    per = f" {ex * 32 / n_part / steps:6.2f}/step" if n_part else ""
    print(f"{a0:06x}: exec {100 * ex / tot_exec:5.1f}%{per}  samples {100 * sm / tot_samp:5.1f}%  [{', '.join(f'{k[6:]} {v}' for k, v in top)}]  {topops}")
