"""Print selected metrics of an `ncu --page raw --csv` dump.  usage: ncu_raw_pick.py raw.csv [substr ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
want = sys.argv[2:] or ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct',
                        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct',
                        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
                        'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct', 'l1tex__throughput.avg.pct',
                        'sm__throughput.avg.pct', 'launch__registers_per_thread', 'launch__occupancy_limit', 'sm__warps_active.avg.pct',
                        'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size', 'sm__inst_executed_pipe_fp64',
                        'smsp__inst_executed_pipe_fp64', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum']
for i, h in enumerate(hdr):
    if any(h.startswith(w) for w in want):
        print("%-90s %-14s %s" % (h, units[i], vals[i]))
