import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
keys=['gpu__time_duration.sum','sm__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__shared_mem_per_block_dynamic']
for r in rows[2:]:
    print('----', r[hdr.index('Kernel Name')][:60])
    for k in keys:
        if k in hdr: print('  ',k, r[hdr.index(k)])
    st={k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):float(r[hdr.index(k)]) for k in hdr if 'issue_stalled' in k and 'per_issue_active' in k and r[hdr.index(k)]}
    print('   stalls:',sorted(st.items(),key=lambda x:-x[1])[:8])
