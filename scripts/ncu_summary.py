"""Print the roofline-relevant metrics of every kernel in an .ncu-rep (run where ncu exists)."""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'smsp__inst_executed.sum',
        'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'sm__inst_executed_pipe_lsu.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h, units = r[0], r[1]
for row in r[2:]:
    d = dict(zip(h, row))
    print('==', d.get('Kernel Name', '?')[:70])
    for k in KEYS:
        if k in d:
            print('   %-75s %s %s' % (k, d[k], units[h.index(k)]))
