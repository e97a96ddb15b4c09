#!/usr/bin/env python3
"""compare_ref.py <dir written by HeadlessLilypad> -- reference-run outputs vs the committed oracle goldens.

Exit status 0 = the C oracle (and therefore the CUDA path, which is tested bit-for-bit against the same files) reproduces
the Java reference bit for bit on BASELINE config 1: parity PINNED.  Any difference is printed with its first location.
TEST INFRASTRUCTURE."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden"
d = Path(sys.argv[1])
bad = 0


def bits_equal(a, b):
    a = np.asarray(a, np.float32) + np.float32(0)      # -0.0 == +0.0
    b = np.asarray(b, np.float32) + np.float32(0)
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


ref = np.fromfile(d / "config1_trace_1000.bin", "<f4").reshape(1000, 34)
ora = np.fromfile(GOLD / "config1_trace_1000.bin", "<f4").reshape(1000, 34)
ok = bits_equal(ref, ora)
if not ok.all():
    s, c = np.argwhere(~ok)[0]
    print(f"trace: {int((~ok).sum())} of {ok.size} values differ; first at solver step {s + 1}, column {c}: "
          f"java {ref[s, c]!r} oracle {ora[s, c]!r}")
    bad += 1
else:
    print("trace: 1000 solver steps x (force, 32 probes) bit-identical")

f = np.fromfile(d / "config1_fields_100.bin", "<f4").reshape(3, 386, 194)
g = np.load(GOLD / "config1_fields_100.npz")
crc = [int(np.bitwise_xor.reduce((a + np.float32(0)).view(np.uint32).ravel())) for a in f]
norms = [float(np.sqrt(np.sum(a.astype(np.float64) ** 2))) for a in f]
spots_ok = all(bits_equal([a[tuple(i)] for i in g["idx"]], g[nm]).all() for a, nm in zip(f, ("ux", "uy", "p")))
if crc != [int(x) for x in g["crc"]] or not spots_ok:
    print(f"fields after 100 steps differ: xor-of-bits {crc} vs {[int(x) for x in g['crc']]}, norms {norms} vs {g['norms'].tolist()}")
    bad += 1
else:
    print("fields after 100 steps: xor-of-bits, spot values and norms identical")

try:
    from oracle import oracle_py as O
    O.build()
    st = O.read_bdimb(ROOT / "rlfluidcontrol_b200" / "data" / "init_state.bdimb")
    env = O.OracleEnv(literal=False)
    env.set_state(st["ux"], st["uy"], st["p"])
    lines = (d / "config1_obs_64.txt").read_text().split("\n")
    k = 0
    while k < min(64, len([l for l in lines if l.strip()])):
        out = env.driver_step(-1.0)
        if out is not None:
            cl, cd = (np.float32(x) for x in lines[k].split())
            if not bits_equal([cl, cd], [out[0], out[1]]).all():
                print(f"observation {k}: java ({cl!r}, {cd!r}) oracle {out}")
                bad += 1
                break
            a1 = np.float32(0.8 * np.sin(2 * np.pi * k / 25.0)); a2 = np.float32(-0.8 * np.sin(2 * np.pi * k / 25.0 + 1.0))
            env.set_xi(a1, a2)
            k += 1
    else:
        print(f"observations: {k} RL steps (Cl, Cd) identical")
except Exception as ex:   # the oracle library is optional for this script
    print("observations not compared:", ex)
sys.exit(1 if bad else 0)
