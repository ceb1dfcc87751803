import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]; vals=rows[2]
want=sys.argv[2:] or ['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_active','smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','l1tex__t_bytes.sum','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__average_warp_latency_per_inst_issued.ratio']
for i,h in enumerate(hdr):
    if h in want or ('warp_issue_stalled' in h and h.endswith('per_warp_active.pct')) or (len(sys.argv)>2 and any(w in h for w in sys.argv[2:])):
        print('%-90s %-14s %s'%(h,units[i],vals[i]))

d={h:v for h,v in zip(hdr,vals)}
items=[]
for h,v in d.items():
    if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued'):
        try: items.append((float(v),h))
        except Exception: pass
tot=sum(x for x,_ in items) or 1.0
print('warp stall sampling (share of samples):')
for x,h in sorted(items,reverse=True)[:8]:
    print('  %-60s %5.1f%%'%(h.replace('smsp__pcsamp_warps_issue_stalled_',''),100*x/tot))
