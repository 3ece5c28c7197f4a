#!/usr/bin/env python
"""Print the handful of ncu raw-page metrics we base decisions on (one block per captured launch)."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum',
        'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__t_sectors_srcunit_tex_op_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__cycles_active.avg',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.sum', 'lts__t_sectors_lookup_miss.sum',
        'lts__d_atomic_input_cycles_active.sum', 'lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed']


def main(path, pattern=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx['Kernel Name']]
        if pattern and pattern not in name:
            continue
        print('-----', name[:80])
        for w in WANT:
            if w in idx:
                print("  %-86s %18s %s" % (w, r[idx[w]][:18], units[idx[w]]))


if __name__ == '__main__':
    main(*sys.argv[1:3])
