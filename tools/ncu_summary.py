"""Compact text summary of an `ncu --set full` report (one block per captured kernel): duration, instruction count,
issue utilisation, occupancy, DRAM / L2 / L1 traffic, stall reasons per issued instruction and, when the report
carries source counters, the opcode mix.  The files under profiles/ are made with it.

  python tools/ncu_summary.py gpurun_out/r02_cl_v6_target.ncu-rep > profiles/r02_roialign_cl_target_ncu.txt"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter

KEYS = [
    ("gpu__time_duration.sum", "duration (us)"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy (%)"),
    ("sm__warps_active.avg.per_cycle_active", "resident warps per SM (avg)"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA (KB)"),
    ("launch__occupancy_limit_registers", "CTAs/SM limit: registers"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM limit: shared memory"),
    ("launch__waves_per_multiprocessor", "waves per SM"),
    ("dram__bytes_read.sum", "DRAM read (MB)"),
    ("dram__bytes_write.sum", "DRAM write (MB)"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> L1 read (MB)"),
    ("l1tex__m_l1tex2xbar_write_bytes.sum", "L1 -> L2 write (MB)"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate (%)"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate (%)"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data pipe busy (%)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput (% of peak)"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe busy (%)"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe busy (%)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = page(rep, "raw")
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"kernel   {d['Kernel Name'][:110]}")
        print(f"launch   grid {d['Grid Size']} block {d['Block Size']}")
        for k, label in KEYS:
            if d.get(k) not in (None, "", "n/a"):
                print(f"  {label:34s} {d[k]}")
        print("  stall cycles per issued instruction:")
        st = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(d[k])
              for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and d[k] not in ("", "n/a")}
        for k, v in sorted(st.items(), key=lambda kv: -kv[1]):
            if v >= 0.05:
                print(f"    {k:22s} {v:6.2f}")
        print()
    src = page(rep, "source")
    hi = next((i for i, r in enumerate(src) if "Instructions Executed" in r), None)
    if hi is not None:
        h = src[hi]
        ia, isrc = h.index("Instructions Executed"), h.index("Source")
        mix = Counter()
        for r in src[hi + 1:]:
            if len(r) <= ia or not r[ia].isdigit():
                continue
            m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", r[isrc])
            if m:
                mix[m.group(1)] += int(r[ia])
        tot = sum(mix.values())
        if tot:
            print(f"opcode mix of the first kernel (warp instructions, total {tot}):")
            for op, n in mix.most_common(14):
                print(f"  {op:10s} {n:10d} {100 * n / tot:5.1f}%")


if __name__ == "__main__":
    main()
