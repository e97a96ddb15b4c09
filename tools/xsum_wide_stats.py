import sys, json
import numpy as np
sys.path.insert(0, '.')
import rlfluidcontrol_b200 as R
res = int(sys.argv[1]); steps = int(sys.argv[2])
t_step = float(np.float32(float(sys.argv[3]) if len(sys.argv) > 3 else 0.18) / np.float32(res))
with R.AFCCylinderBatch(1, init_state=None, resolution=res, x_lengths=16, y_lengths=8, t_step=t_step) as env:
    for k in range(steps):
        env.update2(np.array([[0.5, -0.5]], np.float32) if k == 3 else None)
        if k % 5 == 4 or k == steps - 1:
            st = env.field_sum_stats()[0].tolist()
            print(res, k, "rec", st[0], "walk", st[1] & 0xffff, "hdr-rejected", st[1] >> 16, "ent", st[2] & 0xffff, "key-mismatch walks", st[2] >> 16, "redo segs", st[3], "cycles", st[4:], flush=True)
