"""Key metrics per captured launch of an .ncu-rep (ncu -i ... --page raw --csv), one line per launch."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
H = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.avg.per_cycle_active', 'smsp__inst_executed.sum', 'sm__cycles_active.avg', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum.per_second', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio']
idx = {w: H.index(w) for w in want if w in H}
filt = sys.argv[2] if len(sys.argv) > 2 else ''
for r in rows[2:]:
    if filt and filt not in r[idx['Kernel Name']]:
        continue
    print('----', r[idx['Kernel Name']][:90])
    for w in want[1:]:
        if w in idx:
            print(f'   {w:95s} {r[idx[w]]}')
