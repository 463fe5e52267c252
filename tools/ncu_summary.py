"""Print the metrics that matter for the FP64-bound kernels from an .ncu-rep (developer tool)."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread ',
        'launch__occupancy_limit', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum ',
        'smsp__average_warps_issue_stalled', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct', 'sm__inst_executed_pipe_lsu', 'sm__inst_executed_pipe_xu',
        'sm__inst_executed_pipe_alu.avg.pct', 'sm__inst_executed_pipe_fma.avg.pct', 'smsp__warps_active.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared', 'smsp__inst_executed_op_shared', 'sm__cycles_elapsed.max ',
        'lts__t_sector_hit_rate', 'sm__inst_executed_pipe_uniform', 'smsp__thread_inst_executed_per_inst_executed.ratio']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('==', r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
        for h, u, v in zip(hdr, units, r):
            if any((h + ' ').startswith(k) or k.strip() in h for k in KEYS):
                if 'stalled' in h and ('pcsamp' in h or 'not_issued' in h):
                    continue
                print('  %-90s %-14s %s' % (h, u, v))


if __name__ == '__main__':
    main(sys.argv[1])
