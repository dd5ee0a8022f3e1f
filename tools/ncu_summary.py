import csv, sys, subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
H=rows[0]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','launch__grid_size','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct']
ki=H.index('Kernel Name')
for r in rows[2:]:
    print('==',r[ki][:40])
    for k in keys:
        if k in H: print('   %-75s %s %s'%(k,r[H.index(k)],rows[1][H.index(k)]))
    st=[(float(r[i]),h) for i,h in enumerate(H) if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio') or ('smsp__average_warp_latency_issue_stalled' in h)]
    st=[x for x in st if x[0]>0.3]
    for v,h in sorted(st,reverse=True)[:7]: print('     stall %-60s %.2f'%(h.replace('smsp__average_warps_issue_stalled_','').replace('smsp__average_warp_latency_issue_stalled_',''),v))
