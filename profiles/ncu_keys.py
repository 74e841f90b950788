import csv,sys
r=list(csv.reader(sys.stdin))
hdr=r[0]; units=r[1]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__thread_inst_executed.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__cycles_elapsed.avg','sm__cycles_active.avg','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__cycles_active.avg','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
keys += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
for row in r[2:]:
    print('---', row[hdr.index('Kernel Name')][:60] if 'Kernel Name' in hdr else '')
    for k in keys:
        if k in hdr:
            v=row[hdr.index(k)]
            try:
                if k.startswith('smsp__average_warps_issue_stalled') and float(v)<0.05: continue
            except: pass
            print(f'{k:95s} {v:>18s} {units[hdr.index(k)]}')
