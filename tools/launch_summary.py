"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and share.

  python tools/launch_summary.py gpurun_out/r02_launches_infer.csv [--last-step-of N]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r"^void ", "", r[kn])
        name = re.sub(r"<unnamed>::", "", name)
        name = re.sub(r"[<(].*", "", name).split("::")[-1]
        v = float(r[mv].replace(",", ""))
        unit = r[hdr.index("Metric Unit")]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print(f"{'kernel':48s} {'launches':>8s} {'total us':>10s} {'avg us':>8s} {'share':>6s}")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"{k[:48]:48s} {cnt[k]:8d} {tot[k]:10.1f} {tot[k] / cnt[k]:8.2f} {100 * tot[k] / total:5.1f}%")
    print(f"{'TOTAL':48s} {sum(cnt.values()):8d} {total:10.1f}")


if __name__ == "__main__":
    main()
