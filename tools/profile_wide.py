"""Per-kernel CUDA-event times of a few solver steps of a single wide domain (config 3 and its scaled versions);
diagnostic, not a benchmark.  usage: profile_wide.py [resolution=128] [steps=4] [warm=6]"""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import rlfluidcontrol_b200 as R

res = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 6
t_step = float(np.float32(0.18) / np.float32(res))
with R.AFCCylinderBatch(1, init_state=None, resolution=res, x_lengths=16, y_lengths=8, t_step=t_step) as env:
    for _ in range(warm):
        env.update2()
    env.set_profiling(True)
    its = []
    for _ in range(steps):
        env.update2()
        its.append(env.mg_iters()[0].tolist())
    prof = env.get_profile()
    env.set_profiling(False)
    st = env.field_sum_stats()
tot = sum(r["ms"] for r in prof)
rows = [{"kernel": r["name"], "launches": r["launches"], "avg_us": 1e3 * r["ms"] / max(r["launches"], 1),
         "share": r["ms"] / tot} for r in sorted(prof, key=lambda r: -r["ms"])]
print(json.dumps({"grid": f"{16 * res}x{8 * res}", "steps": steps, "ms_per_step_eager_sum": tot / steps, "mg_iters": its,
                  "field_sum_stats": st[0].tolist(), "kernels": rows}, indent=1))
