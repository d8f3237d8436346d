#!/usr/bin/env python
"""Print selected metrics from `ncu -i X.ncu-rep --page raw --csv` output (stdin or file)."""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
        "sm__warps_active.avg.pct_of_peak", "launch__registers_per_thread", "launch__occupancy_limit",
        "sm__pipe_tensor", "sm__inst_executed_pipe_tensor", "smsp__average_warp", "smsp__average_warps_issue_stalled",
        "sm__throughput.avg.pct", "lts__t_bytes.sum", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct", "sm__maximum_warps_per_active_cycle_pct", "launch__waves_per_multiprocessor",
        "smsp__cycles_active.avg", "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts",
        "smsp__pcsamp_warps_issue_stalled", "sm__cycles_active.avg"]


def main():
    f = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
    rows = list(csv.reader(f))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("==== kernel:", vals[hdr.index("Kernel Name")][:80], "grid", vals[hdr.index("Grid Size")], "block", vals[hdr.index("Block Size")])
        for h, u, v in zip(hdr, units, vals):
            if any(w in h for w in WANT) and v not in ("", "0", "n/a"):
                print(f"{h[:100]:100s} {v:>18s} {u}")


if __name__ == "__main__":
    main()
