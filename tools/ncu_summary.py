#!/usr/bin/env python
"""Extracts the metrics we track from `ncu -i X.ncu-rep --page raw --csv` output.
usage: tools/ncu_summary.py raw.csv [out.csv]"""
import csv
import sys

WANT = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
STALLS = 'smsp__average_warps_issue_stalled_'


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        for i, h in enumerate(hdr):
            if h in WANT or (h.startswith(STALLS) and h.endswith('_per_issue_active.ratio')):
                out.append((h, units[i], vals[i]))
    f = open(sys.argv[2], 'w') if len(sys.argv) > 2 else None
    if f:
        f.write('metric,unit,value\n')
    for h, u, v in out:
        if f:
            f.write('"%s","%s","%s"\n' % (h, u, v))
        try:
            if h.startswith(STALLS) and float(v) < 0.3:
                continue
        except ValueError:
            pass
        print(h.replace(STALLS, 'stall:').replace('_per_issue_active.ratio', ''), u, v[:90])


if __name__ == '__main__':
    main()
