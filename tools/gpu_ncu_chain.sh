#!/bin/bash
# high-rate stall sampling of the chained sweep kernel on the config-3 domain (six consecutive launches = one V-cycle's up leg)
O=gpurun_out/ncu_chain; mkdir -p $O
export RLFC_NO_GRAPH=1
timeout 300 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:k_chain_sweeps3 -s 24 -c 6 -o $O/rep -f python tools/bench_config3.py 6 > $O/run.out 2>&1
ncu -i $O/rep.ncu-rep --page raw --csv > $O/raw.csv 2>/dev/null
ncu -i $O/rep.ncu-rep --page source --csv > $O/src.csv 2>/dev/null
rm -f $O/rep.ncu-rep
python tools/ncu_stalls.py $O/src.csv 40 | cut -c1-150
