#!/bin/bash
# round-2 ncu captures (run on the GPU box through gpurun): launch lists of configs 2 and 3 (eager launches: the graph's
# conditional WHILE node hides the kernels from ncu) and --set full reports of the top kernels.  Outputs: gpurun_out/r02/
O=gpurun_out/r02; mkdir -p $O
export RLFC_NO_GRAPH=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 170 --csv --log-file $O/launches_cfg2.csv python tools/run_steps.py 256 8 > $O/launches_cfg2.out 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file $O/launches_cfg3.csv python tools/bench_config3.py 10 > $O/launches_cfg3.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_mg_coarse_rows|k_smooth0_rows|k_advdif|k_resid_down0_march|k_mg_up0_blk|k_xsum_tables|k_xsum_chain|k_project_shift|k_bc2' -s 28 -c 14 -o $O/full_cfg2 -f python tools/run_steps.py 256 4 > $O/full_cfg2.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_chain_sweeps3|k_chain_incr|k_chain_up|k_chain_down|k_xsum' -s 60 -c 16 -o $O/full_cfg3 -f python tools/bench_config3.py 6 > $O/full_cfg3.out 2>&1
# the reports are too big to travel (64 MiB limit): export the pages here, keep the CSVs only
for c in cfg2 cfg3; do
  ncu -i $O/full_$c.ncu-rep --page raw --csv > $O/full_${c}_raw.csv 2>/dev/null
done
ncu -i $O/full_cfg2.ncu-rep --page source --csv -k regex:k_mg_coarse_rows -c 1 > $O/src_mg_coarse_rows.csv 2>/dev/null
ncu -i $O/full_cfg2.ncu-rep --page source --csv -k regex:k_smooth0_rows -c 1 > $O/src_smooth0_rows.csv 2>/dev/null
ncu -i $O/full_cfg3.ncu-rep --page source --csv -k regex:k_chain_sweeps3 -c 1 > $O/src_chain_sweeps3.csv 2>/dev/null
rm -f $O/*.ncu-rep
ls -la $O
