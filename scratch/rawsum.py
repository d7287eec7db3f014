import csv,sys,subprocess
out=subprocess.run(['ncu','-i',sys.argv[1],'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_elapsed','sm__inst_executed.avg.per_cycle_active','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','launch__waves_per_multiprocessor','smsp__cycles_active.avg','sm__cycles_elapsed.max','sass__inst_executed_local_loads','sass__inst_executed_local_stores','sass__inst_executed_shared_loads','sass__inst_executed_global_loads']
for r in rows[2:]:
    d=dict(zip(hdr,r))
    for k in keys:
        if k in d: print('%-60s %s'%(k, d[k]))
    st={k:float(v.replace(',','')) for k,v in d.items() if 'issue_stalled' in k and 'ratio' in k and v not in ('','n/a')}
    for k,v in sorted(st.items(), key=lambda x:-x[1])[:8]: print('   %-80s %.2f'%(k.replace('smsp__average_warps_issue_stalled_',''),v))
    print('---')
