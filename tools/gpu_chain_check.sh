#!/bin/bash
# parity of the chained smoother + config-3 timing + pipeline stats (run on the GPU box through gpurun)
python -m pytest tests/test_gpu_parity.py -x -q -k "chain or wide_grid" 2>&1 | tail -3
python tools/bench_config3.py 40 > gpurun_out/cfg3_chain.json 2>&1
[ -f rlfluidcontrol_b200/librlfc_stats.so ] && RLFC_LIBRARY=rlfluidcontrol_b200/librlfc_stats.so python tools/bench_config3.py 12 > gpurun_out/chain_stats.log 2>&1
[ -f rlfluidcontrol_b200/librlfc_exp_a.so ] && RLFC_LIBRARY=rlfluidcontrol_b200/librlfc_exp_a.so timeout 120 python tools/bench_config3.py 12 > gpurun_out/chain_exp_a.log 2>&1
true
