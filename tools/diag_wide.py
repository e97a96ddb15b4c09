"""Per-step force / MG iterations / flags of a single wide domain (diagnostic).  usage: diag_wide.py <resolution> <steps> [n_devices]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import rlfluidcontrol_b200 as R

res, steps = int(sys.argv[1]), int(sys.argv[2])
nd = int(sys.argv[3]) if len(sys.argv) > 3 else 1
t_step = float(np.float32(float(sys.argv[4]) if len(sys.argv) > 4 else 0.18) / np.float32(res))
with R.AFCCylinderBatch(1, init_state=None, resolution=res, x_lengths=16, y_lengths=8, t_step=t_step, n_devices=nd) as env:
    for k in range(steps):
        f = env.update2(np.array([[0.5, -0.5]], np.float32) if k == steps // 2 else None)
        print(k, f[0].tolist(), env.mg_iters()[0].tolist(), env.flags().tolist(), flush=True)
        if env.flags().any():
            ux, uy, p = env.get_fields(0)
            for nm, a in (("ux", ux), ("uy", uy), ("p", p)):
                bad = ~np.isfinite(a)
                print(nm, "non-finite:", int(bad.sum()), "first at", (np.argwhere(bad)[0].tolist() if bad.any() else None),
                      "max |finite|", float(np.abs(a[~bad]).max()) if (~bad).any() else None)
            break
