import csv,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
want=["Kernel Name","Grid Size","gpu__time_duration.sum","sm__throughput.avg.pct_of_peak_sustained_elapsed","dram__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__throughput.avg.pct_of_peak_sustained_elapsed","lts__throughput.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","sm__inst_executed.sum","smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","dram__bytes_read.sum","dram__bytes_write.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","memory_l1_wavefronts_shared_ideal","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio","smsp__average_warps_issue_stalled_membar_per_issue_active.ratio","smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio","smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio","smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio","smsp__average_warps_issue_stalled_drain_per_issue_active.ratio","smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio","smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio","smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_selected_per_issue_active.ratio","l1tex__t_sector_hit_rate.pct","sm__cycles_elapsed.max","launch__shared_mem_per_block_dynamic","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.sum"]
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i=hdr.index(w); print(f"{w:86s} {r[i]} {units[i]}")
    print('---')
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
k=0
i=0
while i < len(rows):
    r=rows[i]
    if 'Source' in r and '# Samples' in r:
        hdr=r; si=hdr.index('Source'); sa=hdr.index('# Samples'); ex=hdr.index('Instructions Executed')
        data=[]; i+=1
        while i<len(rows) and not ('Source' in rows[i] and '# Samples' in rows[i]):
            rr=rows[i]
            try: data.append((int(rr[sa]),rr[si],rr[ex]))
            except: pass
            i+=1
        tot=sum(d[0] for d in data)
        print('== kernel',k,'samples',tot,'instrs',len(data)); k+=1
        thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.02
        for j,(v,sx,e) in enumerate(data):
            if v>tot*thr: print(j,v,e,sx[:100])
    else: i+=1
