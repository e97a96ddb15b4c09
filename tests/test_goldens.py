"""Golden-file tests.

* tests/golden/config1_trace_1000.bin, config1_fields_100.npz: frozen OUTPUTS of the oracle (made by
  tests/golden/make_oracle_goldens.py).  The CPU test pins the oracle to them (a silent edit of the C restatement is
  caught), the GPU tests compare the device path with the same files.
* the digest of the reference's own saved/init/init.bdim -- a text checkpoint written by the real Java BDIM.write
  (BDIM.pde:226-237): re-formatting the committed binary fixture with the product's java.lang.Float.toString port must
  reproduce that file byte for byte.  This is a REFERENCE-pinned check: 224 654 floats formatted by the JVM."""
import ctypes as C
import hashlib
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, config1_actions

# sha256 of /root/reference/clientLilypad/saved/init/init.bdim (2 + 386*194 lines, LF line ends)
REFERENCE_INIT_BDIM_SHA256 = "bf76a4ceb86a388be06f3b552c83cbe46766ca78e73046d2b71fc86f6ec22a2f"


def load_trace():
    return np.fromfile(GOLDEN / "config1_trace_1000.bin", np.float32).reshape(1000, 34)


def same(a, b):
    return np.array_equal(np.asarray(a, np.float32), np.asarray(b, np.float32))


def test_oracle_matches_frozen_trace(oracle, init_state):
    """First 160 solver steps (10 RL steps of config 1) of the oracle == the frozen trace, bit for bit."""
    trace = load_trace()
    ref = oracle.OracleEnv(literal=False)
    ref.set_state(init_state["ux"], init_state["uy"], init_state["p"])
    for s in range(160):
        if s % 16 == 0:
            a = config1_actions(s // 16)
            ref.set_xi(a[0], a[1])
        ref.update2()
        assert same(ref.force(), trace[s, :2]), s
        assert same(ref.probes(32), trace[s, 2:]), s


def test_oracle_literal_mode_matches_frozen_trace(oracle, init_state):
    """The literal mode (coefficients rebuilt every step, as the reference does when an action is non-zero,
    BDIM.pde:127) gives the same numbers as the precomputed-geometry mode the goldens were made with."""
    trace = load_trace()
    ref = oracle.OracleEnv(literal=True)
    ref.set_state(init_state["ux"], init_state["uy"], init_state["p"])
    for s in range(32):
        if s % 16 == 0:
            a = config1_actions(s // 16)
            ref.set_xi(a[0], a[1])
        ref.update2()
        assert same(ref.force(), trace[s, :2]), s


def test_checkpoint_text_reproduces_the_reference_file(rlfc):
    """BDIM.write (BDIM.pde:226-237) of the init state, formatted by rlfc_format_float_java, == the file the reference
    ships (written by the JVM), byte for byte."""
    L = rlfc.load_library()
    raw = Path(rlfc.default_init_state()).read_bytes()
    assert raw[:8] == b"RLFCBDIM"
    n, m = np.frombuffer(raw, np.int32, 2, 8)
    t, dt = np.frombuffer(raw, np.float32, 2, 16)
    arr = np.frombuffer(raw, np.float32, 3 * int(n) * int(m), 24).reshape(3, -1)
    L.rlfc_format_float_java.argtypes = [C.c_float, C.c_char_p, C.c_int]
    buf = C.create_string_buffer(64)

    def fmt(v):
        L.rlfc_format_float_java(float(v), buf, 64)
        return buf.value

    cache = {}

    def fmt_cached(v):
        k = v.tobytes()
        s = cache.get(k)
        if s is None:
            s = cache[k] = fmt(v)
        return s

    lines = [fmt(t), fmt(dt)]
    ux, uy, p = arr
    for k in range(ux.size):
        lines.append(fmt_cached(ux[k]) + b", " + fmt_cached(uy[k]) + b", " + fmt_cached(p[k]))
    text = b"\n".join(lines) + b"\n"
    assert len(lines) == 2 + 386 * 194
    assert hashlib.sha256(text).hexdigest() == REFERENCE_INIT_BDIM_SHA256
    # and Float.toString round-trips: parsing the text gives the fixture back bit for bit
    back = np.array([float(x) for ln in lines[2:2 + 2000] for x in ln.split(b",")], np.float32).reshape(-1, 3)
    assert np.array_equal(back[:, 0], ux[:2000]) and np.array_equal(back[:, 2], p[:2000])


@pytest.mark.gpu
def test_gpu_matches_frozen_trace_1000_steps(rlfc):
    """north_star: drag/lift/sensor traces over 1000 steps -- the device path against the frozen file, every float equal
    (force and all 32 probes of every solver step)."""
    trace = load_trace()
    with rlfc.AFCCylinderBatch(1) as env:
        for s in range(1000):
            a = config1_actions(s // 16).reshape(1, 2) if s % 16 == 0 else None
            f, pr = env.update2(a, want_probes=True)
            assert same(f[0], trace[s, :2]), s
            assert same(pr[0], trace[s, 2:]), s


@pytest.mark.gpu
def test_gpu_matches_frozen_fields_100_steps(rlfc):
    g = np.load(GOLDEN / "config1_fields_100.npz")
    with rlfc.AFCCylinderBatch(1) as env:
        for s in range(100):
            env.update2(np.array([[0.5, -0.3]], np.float32) if s == 0 else None)
        fields = env.get_fields(0)
    n2 = lambda a: float(np.sqrt(np.sum(a.astype(np.float64) ** 2)))
    assert [n2(a) for a in fields] == g["norms"].tolist()
    for a, nm, crc, sm in zip(fields, ("ux", "uy", "p"), g["crc"], g["sums"]):
        assert int(np.bitwise_xor.reduce((a + np.float32(0)).view(np.uint32).ravel())) == int(crc), nm
        assert float(a.astype(np.float64).sum()) == float(sm), nm
        for (i, j), v in zip(g["idx"], g[nm]):
            assert a[i, j] == v
