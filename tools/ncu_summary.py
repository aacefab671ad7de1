"""Print the key metrics + warp-stall breakdown of an .ncu-rep (first kernel).  usage: ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys, io
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__block_size', 'launch__grid_size', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_uniform.sum',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_cbu.sum', 'sm__inst_executed_pipe_adu.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__t_bytes.sum', 'sm__cycles_active.avg', 'smsp__cycles_active.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'sm__inst_executed.avg.per_cycle_active', 'smsp__inst_issued.avg.per_cycle_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for vals in rows[2:3]:
    d = dict(zip(hdr, zip(units, vals)))
    for k in KEYS:
        if k in d:
            print('%-70s %-12s %s' % (k, d[k][0], d[k][1]))
    st = []
    for h in hdr:
        if 'smsp__average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try:
                st.append((float(d[h][1].replace(',', '')), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
    print('stalls (warps per issue):', ', '.join('%s=%.2f' % (n, v) for v, n in sorted(st, reverse=True) if v > 0.02))
