"""Extract the roofline-relevant metrics of the first kernel in an .ncu-rep into a small markdown summary."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum.per_second',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'gpc__cycles_elapsed.avg.per_second',
        'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum', 'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_uniform.sum',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem']


def main(rep, title):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units, row = r[0], r[1], r[2]
    print('# %s\n' % title)
    print('source: `%s` (ncu --set full --clock-control none), kernel `%s`\n' % (rep, row[hdr.index('Kernel Name')]))
    print('| metric | value | unit |\n|---|---:|---|')
    for i, h in enumerate(hdr):
        if h in KEYS or 'pipe_tensor_cycles_active_realtime.avg.pct' in h or h.startswith('smsp__average_warps_issue_stalled') and row[i] not in ('0', ''):
            if h.startswith('smsp__average_warps_issue_stalled'):
                try:
                    if float(row[i]) < 0.3:
                        continue
                except ValueError:
                    continue
            print('| `%s` | %s | %s |' % (h, row[i], units[i]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
