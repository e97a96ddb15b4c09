#!/bin/bash
# one --set full capture of ONE kernel of the 256-env default batch (eager launches), exported as CSV pages
# usage: gpu_ncu_one.sh <out name> <kernel regex> [skip=4]
O=gpurun_out/ncu_$1; mkdir -p $O
export RLFC_NO_GRAPH=1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s ${3:-4} -c 1 -o $O/rep -f python tools/run_steps.py 256 3 > $O/run.out 2>&1
ncu -i $O/rep.ncu-rep --page raw --csv > $O/raw.csv 2>/dev/null
ncu -i $O/rep.ncu-rep --page source --csv > $O/src.csv 2>/dev/null
rm -f $O/rep.ncu-rep
python - <<P
import csv
rows=list(csv.reader(open("$O/raw.csv"))); hdr=rows[0]; idx={h:i for i,h in enumerate(hdr)}
want=["gpu__time_duration.sum","l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed","l1tex__data_pipe_lsu_wavefronts.max.pct_of_peak_sustained_elapsed","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","sm__issue_active.avg.pct_of_peak_sustained_elapsed","smsp__issue_active.max.pct_of_peak_sustained_active","smsp__inst_executed.sum","launch__registers_per_thread","sm__warps_active.avg.pct_of_peak_sustained_active","dram__bytes_read.sum","dram__bytes_write.sum","sm__throughput.avg.pct_of_peak_sustained_elapsed","lts__t_sector_hit_rate.pct","l1tex__t_sector_hit_rate.pct","smsp__cycles_active.avg","achieved_occupancy","sm__maximum_warps_per_active_cycle_pct","launch__occupancy_limit_registers","launch__waves_per_multiprocessor"]
for r in rows[2:]:
    print(r[idx["Kernel Name"]][:50], {w: r[idx[w]] for w in want if w in idx})
P
