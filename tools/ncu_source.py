#!/usr/bin/env python
"""Aggregate an ncu source page (cuda,sass) per CUDA source line: instructions executed, stall
samples, shared wavefronts.  usage: tools/ncu_source.py report.ncu-rep [topN]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[hi]
    ci = {n: hdr.index(n) for n in ("Instructions Executed", "# Samples", "L1 Wavefronts Shared",
                                    "L1 Wavefronts Shared Ideal", "Thread Instructions Executed")}
    lines = []
    total = 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or not r[0] or not r[0].isdigit():
            continue  # sass rows have empty line number; the cuda row already aggregates them
        def I(k):
            v = r[ci[k]]
            return int(v) if v.isdigit() else 0
        n = I("Instructions Executed")
        total += n
        lines.append((n, I("# Samples"), I("L1 Wavefronts Shared"),
                      I("L1 Wavefronts Shared Ideal"), r[0], r[1].strip()[:110]))
    print(f"total warp-instructions: {total}")
    print(f"{'inst':>10s} {'%':>5s} {'samples':>8s} {'smem_wf':>9s} {'ideal':>9s}  line  source")
    for n, s, w, wi, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"{n:10d} {100.0 * n / max(total, 1):5.1f} {s:8d} {w:9d} {wi:9d}  {ln:>4s}  {src}")


if __name__ == "__main__":
    main()
