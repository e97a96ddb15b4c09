"""Field.sum serial-pass counters after a few solver steps (and cycle split in RLFC_XS_TIMING builds)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import rlfluidcontrol_b200 as R

n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rng = np.random.default_rng(0)
with R.AFCCylinderBatch(n_envs) as env:
    a = np.clip(rng.normal(0, 0.5, (n_envs, 2)), -1, 1).astype(np.float32)
    for k in range(3):
        env.update2(a if k == 0 else None)
    st = env.field_sum_stats()
    print("mean over envs [batches by record, batches redone, entries applied, segments redone | cycles: setup, entries, serial segments, walks]:")
    print(np.round(st.mean(axis=0), 1).tolist())
    print("max:", st.max(axis=0).tolist())
