#!/usr/bin/env python3
"""Print the headline metrics of every kernel in an ncu report. Usage: tools/ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct"]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:90])
        for i, h in enumerate(hdr):
            if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and "average_warps" in h):
                print(f"   {h:95s} {units[i]:12s} {r[i]}")


if __name__ == "__main__":
    main()
