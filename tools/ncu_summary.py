#!/usr/bin/env python
"""Summarise an .ncu-rep: per kernel the counters the design decisions rest on.
usage: python tools/ncu_summary.py file.ncu-rep [kernel-regex]"""
import csv
import re
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem(blocks)"),
    ("launch__occupancy_limit_registers", "occ_lim_regs(blocks)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__inst_executed_op_shared_atom.sum", "smem atomics"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global ld sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global ld requests"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long_scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall short_scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "stall mio_throttle"),
    ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "stall lg_throttle"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall barrier"),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall wait"),
    ("smsp__average_warp_latency_issue_stalled_branch_resolving.ratio", "stall branch"),
    ("smsp__average_warp_latency_issue_stalled_not_selected.ratio", "stall not_selected"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall math_throttle"),
    ("smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio", "stall dispatch"),
    ("smsp__average_warp_latency_issue_stalled_membar.ratio", "stall membar"),
    ("smsp__average_warp_latency_issue_stalled_no_instruction.ratio", "stall no_instruction"),
    ("smsp__average_warp_latency_issue_stalled_imc_miss.ratio", "stall imc_miss"),
]


def main():
    rep = sys.argv[1]
    rx = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in data:
        name = r[idx["Kernel Name"]]
        if rx and not rx.search(name):
            continue
        print("=" * 4, name[:110])
        for key, label in WANT:
            if key in idx:
                print(f"    {label:28s} {r[idx[key]]:>18s} {units[idx[key]]}")


if __name__ == "__main__":
    main()
