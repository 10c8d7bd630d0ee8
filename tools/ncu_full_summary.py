"""One line of key metrics per captured launch of an `ncu --set full` report (the format of profiles/*_ncu_full_*.txt) plus the stall
breakdown; also prints dram bytes per launch as the `traffic` figure bench.py reads from profiles/ncu_traffic.json."""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
H, U = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum']
stalls = [h for h in H if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
ki = H.index('Kernel Name')
for r in rows[2:]:
    parts = [f'Kernel Name={r[ki]}']
    for k in keys:
        if k in H:
            i = H.index(k)
            parts.append(f'{k}={r[i]} {U[i]}'.strip())
    st = sorted(((float(r[H.index(s)].replace(",", "") or 0), s.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for s in stalls), reverse=True)[:6]
    parts.append('stalls per issue: ' + ', '.join(f'{n} {v:.2f}' for v, n in st))
    print(' | '.join(parts))
