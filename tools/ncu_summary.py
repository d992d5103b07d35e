"""Print a compact summary of an .ncu-rep (first kernel): python tools/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in keys:
        if k in d:
            print(f'{k:75s} {d[k]} {units[hdr.index(k)]}')
    print('-- stall reasons (warps per issue-active cycle)')
    st = [(float(v), h.split('issue_stalled_')[1].split('_per_issue')[0]) for h, v in d.items()
          if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
    for v, n in sorted(st, reverse=True)[:8]:
        print(f'   {n:30s} {v:.3f}')
