import sys,csv
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','sm__cycles_elapsed.avg.per_second','dram__bytes_read.sum','dram__bytes_write.sum','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__grid_size']
for r in rows[2:]:
    for w in want:
        for i,h in enumerate(hdr):
            if h==w: print(f"{w:90s} {rows[1][i]:12s} {r[i]}")
    print('---- stalls')
    for i,h in enumerate(hdr):
        if 'average_warp' in h and 'issue_stalled' in h and h.endswith('.ratio'):
            try:
                v=float(r[i])
                if v>0.3: print(f"   {h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','').replace('smsp__average_warp_latency_issue_stalled_',''):40s} {v:.2f}")
            except: pass
