"""Field.sum (Field.pde:311-318) re-phrased as parallel segment summaries (rlfluidcontrol_b200/csrc/exact_sum.cuh).

CPU part: the host/device core of the summaries is compiled for the host (tests/xsum_host.cpp) and must reproduce the
plain serial float loop BIT FOR BIT on arbitrary data, including wrong predictions of the accumulator.
GPU part: rlfc_env_field_sum (the kernels the projection uses) against the oracle's Field.sum on adversarial fields."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libxsum_host.so"
FP = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def xs():
    src = HERE / "xsum_host.cpp"
    hdr = HERE.parent / "rlfluidcontrol_b200" / "csrc" / "exact_sum.cuh"
    if not LIB.exists() or LIB.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        LIB.parent.mkdir(exist_ok=True)
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", str(LIB), str(src)],
                       check=True)
    L = C.CDLL(str(LIB))
    L.xs_serial.restype = C.c_float
    L.xs_serial.argtypes = [FP, C.c_long]
    L.xs_parallel.restype = C.c_float
    L.xs_parallel.argtypes = [FP, C.c_long, C.c_double, C.POINTER(C.c_long)]
    L.xs_tree_vs_sequential.restype = C.c_long
    L.xs_tree_vs_sequential.argtypes = [FP, C.c_long]
    return L


def both(L, a, noise=0.0):
    a = np.ascontiguousarray(a, np.float32)
    st = (C.c_long * 6)()
    s = np.float32(L.xs_serial(a.ctypes.data_as(FP), a.size))
    p = np.float32(L.xs_parallel(a.ctypes.data_as(FP), a.size, noise, st))
    same = s.tobytes() == p.tobytes() or (np.isnan(s) and np.isnan(p))
    return same, s, p, list(st)


def adversarial(kind, n, rng):
    """Families of addends that stress one aspect each of the integer model."""
    if kind == 0:
        return rng.standard_normal(n)
    if kind == 1:
        return rng.standard_normal(n) + 0.3                                  # drifting accumulator: binade changes
    if kind == 2:
        return rng.standard_normal(n) * 10.0 ** rng.integers(-20, 20, n)      # wild magnitudes
    if kind == 3:
        return np.round(rng.standard_normal(n) * 8) / 8                       # few significand bits: exact ties
    if kind == 4:
        return np.where(rng.random(n) < 0.3, 0.0, rng.standard_normal(n)) - 0.5   # zeros, negative accumulator
    if kind == 5:
        return rng.integers(-3, 4, n) * 2.0 ** rng.integers(-30, 3, n)        # powers of two: ties and exact sums
    if kind == 6:
        return np.abs(rng.standard_normal(n)) * 2.0 ** rng.integers(-140, -100, n)   # denormal range
    if kind == 7:
        return rng.standard_normal(n) + 100 * np.sin(np.arange(n) / 50.0)     # oscillating accumulator: sign changes
    if kind == 8:                                                             # accumulator parked on a power of two
        a = rng.standard_normal(n) * 1e-3
        a[0] = 1024.0
        return a
    if kind == 9:                                                             # half-ulp addends against a big accumulator
        a = np.full(n, 2.0 ** -14)
        a[0] = 512.0
        a[rng.integers(1, max(2, n), n // 7)] *= -1
        return a
    if kind == 11:                                                            # +1, -1, +1, ...: the accumulator changes binade at
        return np.where(np.arange(n) % 2 == 0, 1.0, -1.0)                     # every addition, whole batches of serial segments
    if kind == 12:                                                            # long cancelling stretch inside an ordinary field
        a = rng.standard_normal(n) * 0.01
        k0, ln = n // 3, min(5000, n // 3)
        a[k0:k0 + ln] = np.where(np.arange(ln) % 2 == 0, 0.75, -0.75)
        a[k0 - 1] = -np.sum(a[:k0 - 1])
        return a
    a = rng.standard_normal(n)                                                # Inf / NaN somewhere
    a[rng.integers(0, n)] = [np.inf, -np.inf, np.nan][int(rng.integers(0, 3))]
    return a


def test_matches_serial_loop_on_adversarial_data(xs):
    rng = np.random.default_rng(20261017)
    rejected = crossed = walked = 0
    for trial in range(1100):
        n = int(rng.integers(1, 6000))
        a = adversarial(trial % 11, n, rng)
        for noise in (0.0, 1e-4, 0.3):        # relative error injected into the predicted accumulator
            same, s, p, st = both(xs, a, noise)
            assert same, (trial, trial % 11, n, noise, s, p, st)
            rejected += st[3]
            crossed += st[4]
            walked += st[5]
    assert rejected > 0, "the validity check of the summaries was never exercised"
    assert crossed > 0 and walked > 0, "batch records / the redo of a whole batch were never exercised"


def test_pressure_fields_of_the_oracle(xs, oracle, init_state):
    """Real data: Field.sum of the pressure after each of 12 solver steps; nearly all segments are summarised."""
    env = oracle.OracleEnv(literal=False)
    env.set_state(init_state["ux"], init_state["uy"], init_state["p"])
    env.set_xi(0.5, -0.3)
    for k in range(12):
        _, _, p = env.get_state()
        a = p[1:-1, 1:-1].ravel()
        same, s, par, st = both(xs, a)
        assert same and s == np.float32(oracle.Field(p.shape[0], p.shape[1], values=p).sum())
        assert st[2] < 0.05 * sum(st[:3]) and st[5] <= 4, st
        env.update2()


def test_record_scan_order_is_immaterial(xs, init_state):
    """The kernel condenses 32 summaries with a tree of table compositions, the CPU model composes left to right:
    every group table must come out bit-identical (composition incl. the dead-half normalisation is associative)."""
    rng = np.random.default_rng(11)
    arrays = [init_state["p"][1:-1, 1:-1].ravel()]
    arrays += [adversarial(k % 11, int(rng.integers(40, 5000)), rng) for k in range(330)]
    for a in arrays:
        a = np.ascontiguousarray(a, np.float32)
        assert xs.xs_tree_vs_sequential(a.ctypes.data_as(FP), a.size) == 0


def test_large_array(xs):
    rng = np.random.default_rng(7)
    a = rng.standard_normal(2048 * 1024) * 0.5 + 0.01
    same, s, p, st = both(xs, a)
    assert same, (s, p, st)


@pytest.mark.gpu
@pytest.mark.parametrize("resolution,xl,yl", [(24, 16, 8), (12, 8, 4), (4, 9, 7)])
def test_device_field_sum(rlfc, oracle, resolution, xl, yl):
    """rlfc_env_field_sum on adversarial pressure fields == the oracle's Field.sum, bit for bit (also with the plain
    serial chain, RLFC_PSUM=serial, through test_alternative_execution_paths)."""
    rng = np.random.default_rng(5)
    B = 13
    with rlfc.AFCCylinderBatch(B, init_state=None, resolution=resolution, x_lengths=xl, y_lengths=yl) as env:
        n, m = env.n, env.m
        fields = []
        for e in range(B):
            p = adversarial(e % 13, n * m, rng).astype(np.float32).reshape(n, m)
            fields.append(p)
            env.set_fields(e, None, None, p)
        got = env.field_sum()
        for e in range(B):
            want = np.float32(oracle.Field(n, m, values=fields[e]).sum())
            assert got[e].tobytes() == want.tobytes() or (np.isnan(got[e]) and np.isnan(want)), (e, got[e], want)


@pytest.mark.gpu
def test_device_field_sum_large_domain(rlfc, oracle):
    """The large-domain path (record blocks condensed in parallel, batches that miss their prediction redone with tables
    built for the true accumulator): 2048x1024 fields with the accumulator hovering around binade boundaries, crossing
    zero, meeting large and tiny addends == the oracle's serial float loop, bit for bit."""
    rng = np.random.default_rng(11)
    res = 128
    with rlfc.AFCCylinderBatch(1, init_state=None, resolution=res, x_lengths=16, y_lengths=8,
                               t_step=float(np.float32(0.18) / np.float32(res))) as env:
        n, m = env.n, env.m
        N = n * m
        for case in range(10):
            if case < 5:
                p = adversarial(case * 2 + 1, N, rng).astype(np.float32)
            elif case >= 8:
                p = adversarial(case + 3, N, rng).astype(np.float32)
            elif case == 5:      # zero-mean smooth field: the sum wanders around 0 and the binade boundaries near it
                x = np.linspace(0, 40 * np.pi, N)
                p = (1e-3 * np.sin(x) + 1e-5 * rng.standard_normal(N)).astype(np.float32)
            elif case == 6:      # accumulator parked next to 2.0 by a constant offset, then tiny signed addends
                p = (1e-6 * rng.standard_normal(N)).astype(np.float32)
                p[m + 1] = 2.0
            else:                # long stretches of one sign followed by cancellation
                p = np.where((np.arange(N) // 50000) % 2 == 0, 3e-4, -3e-4).astype(np.float32) * (1 + 1e-3 * rng.standard_normal(N).astype(np.float32))
            p = p.reshape(n, m)
            env.set_fields(0, None, None, p)
            got = env.field_sum()[0]
            want = np.float32(oracle.Field(n, m, values=p).sum())
            assert got.tobytes() == want.tobytes() or (np.isnan(got) and np.isnan(want)), (case, got, want, env.field_sum_stats()[0].tolist())
