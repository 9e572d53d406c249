"""dev tool: per-kernel summary table of an `ncu --page raw --csv` export."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_sector_hit_rate.pct',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed']
want += sorted(h for h in hdr if 'issue_stalled' in h and h.endswith('_per_warp_active.pct') and 'not_issued' not in h)
for w in want:
    if w not in hdr:
        continue
    i = hdr.index(w)
    print('%-72s %-9s' % (w[:72], units[i][:9]), ' | '.join('%12s' % r[i][:12] for r in rows[2:]))
