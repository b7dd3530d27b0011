#!/usr/bin/env python
"""Print selected metrics from an .ncu-rep (raw page) — helper for reading profiles offline.
usage: tools/ncu_metrics.py report.ncu-rep [substring ...]"""
import csv
import subprocess
import sys

DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
           "gpu__dram_throughput.avg.pct", "sm__throughput.avg.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct", "smsp__inst_executed.sum ", "smsp__issue_active.avg.pct",
           "sm__warps_active.avg.pct", "launch__registers_per_thread ", "launch__occupancy_limit",
           "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum ",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum ",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct", "smsp__average_warp", "smsp__average_warps_issue_stalled",
           "lts__t_bytes.sum ", "lts__t_sectors_srcunit_tex_op_read.sum ", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ",
           "sm__cycles_elapsed.avg ", "smsp__cycles_active.avg "]


def main():
    rep = sys.argv[1]
    want = sys.argv[2:] or DEFAULT
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("==", name[:100])
        for h, u, v in zip(hdr, units, r):
            if any(w.strip() in h + " " if w.endswith(" ") else w in h for w in want):
                print(f"  {h:90s} {v} {u}")


if __name__ == "__main__":
    main()
