"""Tiny workload for compute-sanitizer (2 envs, 2 solver steps + 1 RL step on the default grid, or a small grid with
--small); the execution path is selected with the usual RLFC_* environment variables."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import rlfluidcontrol_b200 as R

small = "--small" in sys.argv
kw = dict(resolution=8, x_lengths=16, y_lengths=8, init_state=None) if small else {}
with R.AFCCylinderBatch(2, init_time=-1.0, substeps=2, **kw) as env:   # (2 solver steps per RL step: the sanitizers are slow)
    a = np.array([[0.5, -0.3], [0.0, 0.2]], np.float32)
    for k in range(2):
        f = env.update2(a if k == 0 else None)
    obs, rew, done = env.step(a)
    s = env.field_sum()
    print("ok", f[0].tolist(), obs[0].tolist(), s.tolist(), env.mg_iters().tolist())
