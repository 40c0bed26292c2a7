"""Summarise an `ncu --page raw --csv` dump: duration, pipe utilisation, DRAM bytes, stall reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_uniform.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print(r[hdr.index('Kernel Name')][:90])
    for k in keys:
        if k in hdr:
            print('    %-70s %s %s' % (k, r[hdr.index(k)], rows[1][hdr.index(k)]))
    out = []
    for i, h in enumerate(hdr):
        if 'warps_issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try:
                v = float(r[i])
                if v > 0.1:
                    out.append((v, h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
    print('    stalls (warps per issue-active cycle):', ', '.join('%s=%.2f' % (h, v) for v, h in sorted(out, reverse=True)))
