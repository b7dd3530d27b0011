#!/usr/bin/env python
"""Executed-instruction mix by SASS opcode from an ncu source page. usage: tools/ncu_sass_mix.py rep [topN]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and "Instructions Executed" in r)
hdr = rows[hi]; ie = hdr.index("Instructions Executed"); si = hdr.index("Source")
mix = collections.Counter(); total = 0
for r in rows[hi + 1:]:
    if len(r) <= ie or not r[ie].isdigit(): continue
    toks = r[si].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    n = int(r[ie]); mix[op] += n; total += n
print("total", total)
for op, n in mix.most_common(top): print(f"{n:12d} {100*n/total:5.1f}%  {op}")
