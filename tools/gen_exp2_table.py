#!/usr/bin/env python3
"""Generates include/cpprob/math/exp2_table.inc: T[j] = 2^(j/N), correctly rounded (N = 4096), for dm::exp_weight_tab."""
import os
import sys

import mpmath as mp

mp.mp.dps = 50
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "include", "cpprob", "math", "exp2_table.inc")
N = 4096
rows = [float(mp.power(2, mp.mpf(j) / N)).hex() for j in range(N)]
inc = "// generated (tools/gen_exp2_table.py) — do not edit.  T[j] = 2^(j/4096) rounded to nearest, j = 0..4095\n#define CPPROB_EXP2_TABLE_ROWS \\\n"
for i in range(0, N, 4):
    inc += "    " + ", ".join(rows[i:i + 4]) + ("," if i + 4 < N else "") + " \\\n"
inc += "\n"
open(OUT, "w").write(inc)
