import csv, sys, subprocess
rep = sys.argv[1]
raw = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],stdout=subprocess.PIPE,stderr=subprocess.DEVNULL).stdout.decode()
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
idx={h:i for i,h in enumerate(hdr)}
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','lts__t_bytes.sum','l1tex__t_bytes.sum','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','lts__t_sectors.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
stall=[h for h in hdr if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('_per_issue_active.ratio')] or [h for h in hdr if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct')]
for r in rows[2:]:
    print('=====', r[idx['Kernel Name']][:70])
    for w in want:
        if w in idx: print('  %-70s %s %s'%(w, r[idx[w]][:40], units[idx[w]]))
    st=[(float(r[idx[h]]) if r[idx[h]] not in ('','nan','-nan') else 0.0,h) for h in stall]
    for v,h in sorted(st,reverse=True)[:7]: print('   stall %-80s %.2f'%(h.replace('smsp__average_warps_issue_stalled_','').replace('smsp__average_warp_latency_issue_stalled_',''),v))
