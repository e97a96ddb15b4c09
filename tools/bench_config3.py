"""BASELINE config 3: one 2048x1024 cylinder-wake domain on one B200 (resolution 128, 16x8 lengths,
t_step = 0.18/128, uniform start, actions 0 then (0.5, -0.5)).  Reports solver-steps/s and MG iterations per solve.
A functional measurement of the wide-grid fallback path, not the headline benchmark (bench.py)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import rlfluidcontrol_b200 as R

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
res = int(sys.argv[2]) if len(sys.argv) > 2 else 128
t_step = float(np.float32(0.18) / np.float32(res))
with R.AFCCylinderBatch(1, init_state=None, resolution=res, x_lengths=16, y_lengths=8, t_step=t_step) as env:
    env.update2()                                     # graph build + warm-up
    its = []
    t0 = time.perf_counter()
    for k in range(steps):
        env.update2(np.array([[0.5, -0.5]], np.float32) if k == steps // 2 else None)
        its.append(env.mg_iters()[0].tolist())
    dt = time.perf_counter() - t0
    print(json.dumps({"config": f"single {16 * res}x{8 * res} domain, uniform start", "solver_steps_per_s": steps / dt,
                      "ms_per_solver_step": 1e3 * dt / steps, "mg_iters_per_solve_mean": float(np.mean(its)),
                      "mg_iters_max": int(np.max(its)), "steps": steps}))
