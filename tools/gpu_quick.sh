#!/bin/bash
# quick GPU check: parity subset + config-2 bench with the per-kernel table (run on the GPU box through gpurun)
# usage: gpu_quick.sh <out dir name> [library variants ...]
O=gpurun_out/$1; shift; mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_protocol_and_scale.py -m gpu -x -q 2>&1 | tail -4 > $O/pytest.log
cat $O/pytest.log
for v in default "$@"; do
  if [ "$v" = default ]; then unset RLFC_LIBRARY; else export RLFC_LIBRARY=rlfluidcontrol_b200/librlfc_$v.so; fi
  python bench.py --no-cpu-baseline --steps 6 > $O/bench_cfg2_$v.json 2> $O/bench_cfg2_$v.err
  python - <<P
import json
d=json.load(open("$O/bench_cfg2_$v.json"))
print("$v", round(d["value"],1), round(d["e2e"]["value"],1), " ".join(f'{k["kernel"]}={k["avg_ms"]*1e3:.0f}' for k in d["kernels"][:9]))
P
done
