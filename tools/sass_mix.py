#!/usr/bin/env python3
"""Instruction mix of one kernel's SASS (static count), grouped by issue pipe.

usage: tools/sass_mix.py <lib.so|.o|.cubin> <kernel-name-substring> [--dump] [--range 0xLO 0xHI]
The FP64 pipe share of the particle loop is the static proxy for the ncu metric
sm__inst_executed_pipe_fp64 ("K1 instruction budget").
"""
import collections
import re
import subprocess
import sys


def kernel_sass(path, needle):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    blocks = re.split(r"\n\s*Function : ", out)
    for b in blocks[1:]:
        name = b.split("\n", 1)[0].strip()
        if needle in name:
            return name, b
    raise SystemExit(f"no kernel matching {needle!r}")


def classify(op):
    base = op.split(".")[0]
    if base in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"):
        return "fp64"
    if base in ("MUFU",):
        return "xu"
    if base in ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IMUL"):
        return "fma"
    if base in ("LDG", "STG", "LDS", "STS", "LDC", "LDL", "STL", "ATOMG", "ATOMS", "RED", "LDCU"):
        return "lsu"
    if base in ("BRA", "EXIT", "BAR", "BSSY", "BSYNC", "CALL", "RET", "WARPSYNC", "NOP", "BREAK", "YIELD"):
        return "ctrl"
    if base in ("F2F", "I2F", "F2I", "I2I", "FRND", "DSETP"):
        return "conv"
    return "alu"


BRANCH = re.compile(r"\bBRA(?:\.\w+)*\s+(?:!?\w+,\s*)*0x([0-9a-f]+)")   # BRA 0x..., @P0 BRA P3, 0x..., BRA.U !UP0, 0x...
MARKERS = ("LDS", "CALL.REL.NOINC", "MUFU.RSQ64H")   # what only the steady-state particle loop issues most of: the table
# loads of the ziggurat / exp (current build), the slow-path call sites (first ziggurat build), Box-Muller's sqrt seed


def loop_marker(body):
    return next((m for m in MARKERS if m in body), MARKERS[0])


def particle_loop(body, marker=None):
    """(lo, hi) addresses of the innermost loop (backward branch span) with the most `marker` instructions:
    the steady-state particle loop of the SIS kernels."""
    marker = marker or loop_marker(body)
    instr = []
    for line in body.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            instr.append((int(m.group(1), 16), m.group(2)))
    marks = [a for a, t in instr if marker in t]
    best, best_key = None, None
    for a, t in instr:
        m = BRANCH.search(t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        n_marks = sum(1 for x in marks if tgt <= x <= a)
        if tgt < a and n_marks:
            # reject spans that contain another backward branch (outer loops)
            inner_back = any(tgt < a2 < a and (mm := BRANCH.search(t2)) and int(mm.group(1), 16) < a2
                             and int(mm.group(1), 16) >= tgt for a2, t2 in instr)
            if inner_back:
                continue
            key = (n_marks, -(a - tgt))
            if best_key is None or key > best_key:
                best, best_key = (tgt, a), key
    return best


def loop_budget(path, needle):
    """opcode counts of the particle loop: dict(total=, fp64=, dfma=, dadd=, dmul=, dsetp=)"""
    name, body = kernel_sass(path, needle)
    lo, hi = particle_loop(body)
    c = collections.Counter()
    for line in body.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and lo <= int(m.group(1), 16) <= hi:
            c["total"] += 1
            base = m.group(2).split(".")[0]
            if m.group(2).startswith("MUFU.RSQ64H"):
                c["box_muller"] += 1            # one per stream pair = 2 particles
            if m.group(2).startswith("CALL.REL.NOINC"):
                c["calls"] += 1
            if m.group(2).startswith("ISETP.GE.U32"):
                c["zig_tests"] += 1             # ziggurat: one high-word fast test per draw = per particle of the README model
            if classify(m.group(2)) == "fp64":
                c["fp64"] += 1
            if base in ("DFMA", "DADD", "DMUL", "DSETP"):
                c[base.lower()] += 1
    # instructions that only run on the slow path of a draw (argument set-up between the fast-accept branch and
    # the call): static, but executed by 0.06 % of the draws
    lines = [l for l in body.splitlines() if (mm := re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)) and lo <= int(mm.group(1), 16) <= hi]
    cold = 0
    for i, l in enumerate(lines):
        if "CALL.REL.NOINC" in l:
            j = i
            while j > 0 and "BRA" not in lines[j - 1]:
                j -= 1
            cold += i - j + 2                   # set-up, the call and the jump over the fast tail
    c["cold"] = cold
    c["particles_per_trip"] = c["zig_tests"] if c["zig_tests"] else 2 * c["box_muller"]
    return dict(c)


def main():
    path, needle = sys.argv[1], sys.argv[2]
    name, body = kernel_sass(path, needle)
    if "--loop" in sys.argv:
        lo, hi = particle_loop(body)
        sys.argv += ["--range", hex(lo), hex(hi)]
        print(f"particle loop {hex(lo)}..{hex(hi)}")
    ops = collections.Counter()
    classes = collections.Counter()
    lines = []
    lo_addr, hi_addr = 0, 1 << 62
    if "--range" in sys.argv:
        i = sys.argv.index("--range")
        lo_addr, hi_addr = int(sys.argv[i + 1], 16), int(sys.argv[i + 2], 16)
    for line in body.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        if not (lo_addr <= int(m.group(1), 16) <= hi_addr):
            continue
        op = m.group(2)
        lines.append(line.rstrip())
        ops[op.split(".")[0]] += 1
        classes[classify(op)] += 1
    total = sum(classes.values())
    print(name)
    print(f"total {total}")
    for k, v in classes.most_common():
        print(f"  {k:5s} {v:5d}  {100.0 * v / total:5.1f}%")
    print("top opcodes:", ", ".join(f"{k}:{v}" for k, v in ops.most_common(24)))
    if "--dump" in sys.argv:
        print("\n".join(lines))


if __name__ == "__main__":
    main()
