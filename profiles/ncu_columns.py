import csv,sys,subprocess
rep=sys.argv[1]; step=int(sys.argv[2]) if len(sys.argv)>2 else 1
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; data=rows[2:]
want=["Kernel Name","gpu__time_duration.sum","smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__grid_size",
"l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed","l1tex__t_sector_hit_rate.pct","lts__t_sector_hit_rate.pct","dram__bytes_read.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
"smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio","smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio","smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
"sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active","sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active","sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fmaheavy.sum.pct_of_peak_sustained_active","sm__inst_executed_pipe_fmalite","sm__inst_executed_pipe_uniform.sum.pct","sass__inst_executed_local_loads"]
idx={h:i for i,h in enumerate(hdr)}
for w in want:
    if w not in idx:
        c=[h for h in hdr if w in h]
        if not c: continue
        w=c[0]
    vals=[r[idx[w]] for r in data[::step]]
    if w=="Kernel Name": vals=[v.replace('void cacao::<unnamed>::','')[:24] for v in vals]
    else:
        try: vals=[f"{float(v.replace(',','')):.4g}" for v in vals]
        except: pass
    print(f"{w[:58]:58s}", *[f"{v:>12s}" for v in vals])
