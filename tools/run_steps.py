"""Run a few solver steps of the default 256-env batch (used under ncu; not a benchmark)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import rlfluidcontrol_b200 as R

n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(0)
with R.AFCCylinderBatch(n_envs) as env:
    a = np.clip(rng.normal(0, 0.5, (n_envs, 2)), -1, 1).astype(np.float32)
    for k in range(steps):
        f = env.update2(a if k == 0 else None)
    print("ok", f[0], env.launch_count)
