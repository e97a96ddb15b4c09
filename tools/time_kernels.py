"""Per-kernel CUDA-event times of a few solver steps (eager launches); diagnostic, not a benchmark."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import rlfluidcontrol_b200 as R

n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rng = np.random.default_rng(0)
with R.AFCCylinderBatch(n_envs) as env:
    a = np.clip(rng.normal(0, 0.5, (n_envs, 2)), -1, 1).astype(np.float32)
    env.update2(a)
    env.set_profiling(True)
    for _ in range(steps):
        env.update2()
    prof = env.get_profile()
    env.set_profiling(False)
print(" ".join(f"{r['name']}={r['ms'] / max(r['launches'], 1):.4f}" for r in sorted(prof, key=lambda r: -r["ms"])))
