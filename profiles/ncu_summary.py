#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump: per kernel duration, DRAM traffic, occupancy, stall mix."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]
idx={h:i for i,h in enumerate(hdr)}
pat=sys.argv[2] if len(sys.argv)>2 else ''
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','smsp__inst_executed.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__warps_eligible.avg.per_cycle_active','lts__t_bytes.sum','sm__maximum_warps_per_active_cycle_pct']
for r in rows[2:]:
    name=r[idx['Kernel Name']]
    if pat and pat not in name: continue
    print('----', name[:80])
    for k in keys:
        if k in idx: print('   %-70s %-14s %s'%(k,units[idx[k]],r[idx[k]]))
    st=[]
    for h in hdr:
        if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
            try: st.append((float(r[idx[h]]),h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
            except: pass
    print('   stalls (warps per issue-active cycle):', ', '.join('%s=%.2f'%(n,v) for v,n in sorted(st,reverse=True)[:7]))
