"""Freeze oracle outputs as golden files (run in the build container; the files are committed):

  config1_trace_1000.bin   BASELINE config 1: from the shipped init state, 1000 solver steps, the action of RL step k
                           (conftest.config1_actions) applied at solver step 16 k; per solver step the raw force
                           (fx, fy) and the 32 surface-pressure probes: float32 [1000][34]
  config1_fields_100.npz   norms and spot values of ux, uy, p after 100 solver steps with actions (0.5, -0.3)

The oracle is the checker, these files pin IT: a later edit of oracle/lilypad_oracle.c that changes any of these
numbers fails tests/test_oracle.py (CPU) and the GPU path is compared with the same files (tests/test_gpu_goldens.py).
They are outputs of the C restatement, not of the Java reference (which cannot run here: no JVM)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import config1_actions  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

O.build()
st = O.read_bdimb(ROOT / "rlfluidcontrol_b200" / "data" / "init_state.bdimb")
HERE = Path(__file__).resolve().parent

ref = O.OracleEnv(literal=False)
ref.set_state(st["ux"], st["uy"], st["p"])
trace = np.zeros((1000, 34), np.float32)
for s in range(1000):
    if s % 16 == 0:
        a = config1_actions(s // 16)
        ref.set_xi(a[0], a[1])
    ref.update2()
    trace[s, :2] = ref.force()
    trace[s, 2:] = ref.probes(32)
trace.tofile(HERE / "config1_trace_1000.bin")

ref = O.OracleEnv(literal=False)
ref.set_state(st["ux"], st["uy"], st["p"])
ref.set_xi(0.5, -0.3)
for s in range(100):
    ref.update2()
ux, uy, p = ref.get_state()
n2 = lambda a: float(np.sqrt(np.sum(a.astype(np.float64) ** 2)))
idx = [(1, 1), (96, 96), (100, 97), (120, 80), (200, 50), (384, 192), (385, 193)]
np.savez(HERE / "config1_fields_100.npz", norms=np.array([n2(ux), n2(uy), n2(p)]),
         idx=np.array(idx), ux=np.array([ux[i] for i in idx]), uy=np.array([uy[i] for i in idx]), p=np.array([p[i] for i in idx]),
         crc=np.array([int(np.bitwise_xor.reduce((a + np.float32(0)).view(np.uint32).ravel())) for a in (ux, uy, p)], dtype=np.uint64),
         sums=np.array([float(a.astype(np.float64).sum()) for a in (ux, uy, p)]))
print("written", HERE / "config1_trace_1000.bin", HERE / "config1_fields_100.npz")
