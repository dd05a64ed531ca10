#!/usr/bin/env python3
"""Prints the handful of `ncu --page raw --csv` columns that matter for this repo's kernels (one line per captured launch)."""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__inst_executed.sum", "warp_inst"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "ipc/SM"),
    ("smsp__issue_active.avg.pct", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "lim_reg"),
    ("launch__occupancy_limit_shared_mem", "lim_smem"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fmaI%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conf"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct", "st_long"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_disp"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"),
    ("smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "st_imc"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0]
        out = [name[:34]]
        seen = set()
        for k, short in KEYS:
            if k in idx and short not in seen:
                seen.add(short)
                out.append(f"{short}={r[idx[k]]}{units[idx[k]] if short in ('time','dram_rd','dram_wr') else ''}")
        print("  ".join(out))


if __name__ == "__main__":
    main(sys.argv[1])
