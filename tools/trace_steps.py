"""Timeline of the kernels of one solver step per env group (eager launches bracketed by CUDA events on the group
streams; RLFC_EAGER_GROUPS=1 RLFC_FIXED_ITERS=1 RLFC_GROUPS=g RLFC_TRACE=<csv>).  Diagnostic only, not a benchmark."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
trace = os.environ.setdefault("RLFC_TRACE", "gpurun_out/trace.csv")
os.environ.setdefault("RLFC_EAGER_GROUPS", "1")
os.environ.setdefault("RLFC_FIXED_ITERS", "1")
import rlfluidcontrol_b200 as R

n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
if os.path.exists(trace):
    os.remove(trace)
rng = np.random.default_rng(0)
with R.AFCCylinderBatch(n_envs) as env:
    a = np.clip(rng.normal(0, 0.5, (n_envs, 2)), -1, 1).astype(np.float32)
    env.update2(a)
    env.update2()
    env.set_profiling(True)
    env.update2(); env.update2()
    env.get_profile()
    env.set_profiling(False)
rows = [ln.strip().split(",") for ln in open(trace)]
rows = [(int(g), k, float(a), float(b)) for g, k, a, b in rows]
t0 = min(r[2] for r in rows)
end = max(r[3] for r in rows)
print(f"groups={len(set(r[0] for r in rows))} span={end - t0:.3f} ms for 2 solver steps")
for g in sorted(set(r[0] for r in rows)):
    print(f"-- group {g}")
    for r in rows:
        if r[0] == g:
            print(f"   {r[1]:12s} {r[2] - t0:8.3f} -> {r[3] - t0:8.3f}  ({r[3] - r[2]:.3f})")
