"""Summarise an .ncu-rep (read here with `ncu -i ... --page raw --csv`): one block of key metrics per captured launch."""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'smsp__cycles_active.avg', 'launch__shared_mem_per_block_dynamic',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum',
        'sm__inst_executed_pipe_lsu.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']
rows = list(csv.reader(subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
hdr = rows[0]
names = [r[hdr.index('Kernel Name')][:60] for r in rows[2:]]
print('kernels:', names)
for k in KEYS:
    if k in hdr:
        i = hdr.index(k); print(f"{k} [{rows[1][i]}]:", [r[i] for r in rows[2:]])
stall = []
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
        stall.append((h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), [float(r[i]) for r in rows[2:]]))
for n in range(len(names)):
    print('stalls', names[n][:40], sorted(((round(v[n], 2), k) for k, v in stall), reverse=True)[:7])
