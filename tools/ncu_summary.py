#!/usr/bin/env python
"""Print the counters this repo tracks from an ncu raw page:  ncu -i rep --page raw --csv | python tools/ncu_summary.py"""
import csv
import sys

KEEP = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__sass_branch_targets_threads_divergent.sum"]


def main():
    rows = list(csv.reader(sys.stdin))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print("==", d.get("Kernel Name", ("", "?"))[1][:90])
        for k in KEEP:
            if k in d:
                print(f"{k} [{d[k][0]}] = {d[k][1]}")
        for k in hdr:
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                v = float(d[k][1] or 0)
                if v >= 0.1:
                    print(f"{k.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', '')} = {v:.3f}")


if __name__ == "__main__":
    main()
