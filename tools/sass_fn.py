#!/usr/bin/env python3
"""Prints the SASS of one kernel of the built library, one instruction per line (`offset mnemonic operands`).
usage: python tools/sass_fn.py <regex on the mangled name> [lib]"""
import re
import subprocess
import sys

lib = sys.argv[2] if len(sys.argv) > 2 else "cpprob_b200/lib/libcpprob_sis.so"
text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
idx = [i for i, l in enumerate(text) if "Function :" in l]
for n, i in enumerate(idx):
    if re.search(sys.argv[1], text[i]):
        end = idx[n + 1] if n + 1 < len(idx) else len(text)
        print("# " + text[i].strip())
        for l in text[i:end]:
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*/\*", l)
            if m:
                print(m.group(1), m.group(2))
        break
