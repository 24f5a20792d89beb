import csv,sys,subprocess
f=sys.argv[1]
out=subprocess.run(["ncu","-i",f,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
d={h:(v,u) for h,u,v in zip(hdr,units,vals)}
keys=["Kernel Name","gpu__time_duration.sum","launch__grid_size","launch__block_size","launch__registers_per_thread","launch__shared_mem_per_block_dynamic","launch__occupancy_limit_shared_mem","launch__occupancy_limit_registers","launch__occupancy_limit_warps","sm__warps_active.avg.pct_of_peak_sustained_active","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","sm__throughput.avg.pct_of_peak_sustained_elapsed","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active","sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active","smsp__warps_eligible.avg.per_cycle_active","smsp__inst_executed.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","lts__t_bytes.sum","smsp__average_warp_latency_issue_stalled_barrier.ratio","smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
for k in keys:
    if k in d: print("%-80s %s %s"%(k,d[k][0],d[k][1]))
# stall reasons
st=[(float(v.replace(',','')),h) for h,(v,u) in d.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v not in ("","n/a")]
for v,h in sorted(st,reverse=True)[:7]: print("  stall %-60s %.2f"%(h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")],v))
