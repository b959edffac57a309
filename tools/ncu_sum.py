"""Key metrics + top stall reasons per kernel of an `ncu --page raw --csv` dump: python tools/ncu_sum.py raw.csv"""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]; data=rows[2:]
want=['gpu__time_duration.sum','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.sum','smsp__average_warp_latency_per_inst_issued.ratio','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__cycles_elapsed.avg','smsp__inst_executed.avg','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__warps_eligible.avg.per_cycle_active','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__thread_inst_executed_per_inst_executed.ratio','lts__t_bytes.sum','l1tex__t_bytes.sum']
idx={h:i for i,h in enumerate(hdr)}
st=[h for h in hdr if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
for d in data:
    print('----', d[idx['Kernel Name']][:70])
    for w in want:
        if w in idx: print('  ',w, d[idx[w]], units[idx[w]])
    vals=sorted(((float(d[idx[h]].replace(',','')) if d[idx[h]] else 0,h) for h in st),reverse=True)[:7]
    print('  stalls',[(round(v,2),h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')) for v,h in vals])
