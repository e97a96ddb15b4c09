"""CPU tests of the oracle (oracle/lilypad_oracle.c) against every anchor the reference offers for this path:
the three docstring known-answers, the invariants of the shipped init.bdim fixture, closed forms of the BDIM
kernels, and the drift-check values recorded by an independent restatement."""
import ctypes as C

import numpy as np
import pytest


def test_body_docstring_distance(oracle):
    """Body.pde:16-28: distance from (3,1) to the triangle (0.5,2.5),(2.75,0.25),(0,0) is sqrt(.5)."""
    L = oracle.lib()
    b = L.ora_body_new(0, 0)
    for x, y in [(0.5, 2.5), (2.75, 0.25), (0, 0)]:
        L.ora_body_add(b, x, y)
    L.ora_body_end(b)
    assert abs(L.ora_body_distance(b, 3.0, 1.0) - np.sqrt(0.5)) < 1e-6
    L.ora_body_free(b)


def test_poisson_docstring(oracle):
    """PoissonMatrix.pde:17-31: A*(i^2 - j^2) vanishes in the interior for unit coefficients."""
    c = oracle.VField(40, 40, 0, 0)
    c.x[...] = 1; c.y[...] = 1; c.setBC()
    A = oracle.Poisson(c)
    x = oracle.Field(40, 40, 0, 1)
    i, j = np.meshgrid(np.arange(40), np.arange(40), indexing="ij")
    x.a[...] = (i * i - j * j).astype(np.float32)
    Ax = A.times(x)
    assert np.abs(Ax.a[2:-2, 2:-2]).max() == 0.0


def test_mg_docstring(oracle):
    """MG.pde:8-26: the solver drives the residual of a divergent patch on a 128^2 box below tol."""
    N = 130
    u = oracle.VField(N, N, 1, -0.5)
    u.x[40:50, 40:75] = 0.0
    c = oracle.VField(N, N, 0, 0)
    c.x[...] = 1; c.y[...] = 1; c.setBC()
    A = oracle.Poisson(c)
    x = oracle.Field(N, N)
    b = u.divergence()
    it, rr, tol = A.solve(x, b, 20.0)
    assert 1 <= it < 20 and rr < tol
    assert abs(tol - 128 * 128 * 1e-8) / tol < 1e-4          # tol = sum over the interior of (1e-4)^2
    Ax = A.times(x)                                          # keep the owner alive while .a is read
    r = b.a - Ax.a
    assert np.abs(r).max() < 1e-3


def test_fixture_invariants(init_state):
    """the shipped init.bdim is a state written by BDIM.write after a corrector step."""
    s = init_state
    assert (s["n"], s["m"]) == (386, 194)
    assert abs(s["t"] - 2400.2554) < 1e-3 and np.float32(s["dt"]) == np.float32(np.float32(0.0075) * 24)
    ux, uy, p = s["ux"], s["uy"], s["p"]
    assert np.all(ux[1, :] == 1.0) and np.all(uy[:, 1] == 0.0) and np.all(uy[:, -1] == 0.0)
    assert np.array_equal(p[0, 1:-1], p[1, 1:-1]) and np.array_equal(p[1:-1, 0], p[1:-1, 1])
    assert abs(p[1:-1, 1:-1].astype(np.float64).mean()) < 1e-6
    div = ux[2:, 1:-1] - ux[1:-1, 1:-1] + uy[1:-1, 2:] - uy[1:-1, 1:-1]
    assert np.sqrt((div.astype(np.float64) ** 2).mean()) < 5e-4


def test_bdim_kernels(oracle):
    """BDIM.pde:199-215 closed forms: mu0 clamps to [0,1], mu0(0)=.5, mu1 is even, vanishes at |d|>=eps."""
    L = oracle.lib()
    eps = 2.0
    assert L.ora_bdim_delta0(-2.0, eps) == 0 and L.ora_bdim_delta0(2.0, eps) == 1 and L.ora_bdim_delta0(0.0, eps) == 0.5
    for d in (0.3, 1.1, 1.9):
        assert abs(L.ora_bdim_delta0(d, eps) + L.ora_bdim_delta0(-d, eps) - 1) < 1e-6
        assert abs(L.ora_bdim_delta1(d, eps) - L.ora_bdim_delta1(-d, eps)) < 1e-7
        ref = 0.25 * (eps - d * d / eps) - 1 / (2 * np.pi) * (d * np.sin(d * np.pi / eps) + eps / np.pi * (1 + np.cos(d * np.pi / eps)))
        assert abs(L.ora_bdim_delta1(d, eps) - ref) < 1e-6
    assert L.ora_bdim_delta1(2.0, eps) == 0 and L.ora_bdim_delta1(-3.0, eps) == 0
    assert L.ora_union_delta0(-1.0) == 0 and L.ora_union_delta0(1.0) == 1


def test_setbc_order(oracle):
    """Field.pde:209-234: ghost gets the OLD boundary-adjacent value, then the face row is overwritten; the
    gradientExit mean correction uses the copied outflow column."""
    rng = np.random.default_rng(0)
    f = oracle.Field(8, 6, btype=1, bval=1.0, gradientExit=True, values=rng.normal(size=(8, 6)))
    old = f.a.copy()
    f.setBC()
    assert np.array_equal(f.a[0, 1:-1], old[1, 1:-1]) and np.all(f.a[1, :] == 1.0)
    s = np.float32(0)
    for j in range(1, 5):
        s = np.float32(s + old[6, j])
    s = np.float32(s / np.float32(4))
    assert np.array_equal(f.a[7, 1:-1], (old[6, 1:-1] + (np.float32(1.0) - s)).astype(np.float32))
    assert f.a[7, 0] == old[6, 1]            # corner copied before the correction


def test_survey_drift_values(oracle, init_state):
    """values recorded by an independently written restatement of the same reference code."""
    e = oracle.OracleEnv(literal=False)
    e.set_state(init_state["ux"], init_state["uy"], init_state["p"])
    got = []
    for _ in range(3):
        e.update2()
        got.append(e.force())
        assert e.mg_iters() == (1, 1)
    expect = [(13.6044025, 0.103305936), (13.6054707, 0.104807034), (13.6061077, 0.10624747)]
    for g, x in zip(got, expect):
        assert g[0] == np.float32(x[0]) and g[1] == np.float32(x[1])
    e2 = oracle.OracleEnv(literal=False)
    e2.set_state(init_state["ux"], init_state["uy"], init_state["p"])
    obs = None
    for _ in range(16):
        obs = e2.driver_step()
    assert abs(float(obs[0]) - 0.009605) < 5e-7 and abs(float(obs[1]) - 1.133926) < 5e-7


def test_action_step_and_norms(oracle, init_state):
    e = oracle.OracleEnv(literal=False)
    e.set_state(init_state["ux"], init_state["uy"], init_state["p"])
    e.set_xi(0.5, -0.3)
    e.update2()
    fx, fy = e.force()
    assert fx == np.float32(13.2472315) and fy == np.float32(0.26071915)
    for _ in range(29):
        e.update2()
    ux, uy, p = e.get_state()
    assert np.isfinite(ux).all() and np.isfinite(p).all()


def test_literal_equals_cached(oracle, init_state):
    """Rebuilding coefficients / Poisson matrices where the reference does (BDIM.pde:127, MG.pde:70) gives the same
    numbers as caching them: the geometry never moves (CircleBody.rotate, Body.pde:412-415)."""
    a = oracle.OracleEnv(literal=True)
    b = oracle.OracleEnv(literal=False)
    for e in (a, b):
        e.set_state(init_state["ux"], init_state["uy"], init_state["p"])
        e.set_xi(0.7, -0.2)
    for _ in range(3):
        a.update2(); b.update2()
    for x, y in zip(a.get_state(), b.get_state()):
        assert np.array_equal(x, y)
    assert a.force() == b.force()


def test_probes_and_linear(oracle, init_state):
    e = oracle.OracleEnv(literal=False)
    e.set_state(init_state["ux"], init_state["uy"], init_state["p"])
    pr = e.probes(32)
    p = init_state["p"]
    # probe 0 sits at (n/4 + D/2, m/2) = (108, 96): integer coordinates -> exact cell value (Field.pde:184-185)
    assert pr[0] == p[108, 96]
    assert np.isfinite(pr).all()


def test_reward_restatement(oracle):
    r = oracle.reference_reward(1.2, (0.5, -1.0))
    assert abs(r - (-1.2 - np.pi / 8 * 0.0097 * 3.66 ** 3 * (0.125 + 1.0))) < 1e-12


def test_adaptive_dt_variant(oracle, init_state):
    """AFCCylinder.update() (AFCCylinder.pde:63-84): dt = min(u.CFL(nu), 1) before every step (BDIM.pde:217-219,
    VectorField.pde:225-235).  On the developed wake max(|ux|+|uy|) is about 1.8, so dt is about 0.51 grid units --
    larger than the fixed 0.18 -- and t advances by dt/resolution."""
    e = oracle.OracleEnv(literal=False)
    e.set_state(init_state["ux"], init_state["uy"], init_state["p"])
    ux, uy = init_state["ux"], init_state["uy"]
    b = max(np.float32(abs(ux[0, 0]) + abs(uy[0, 0])), (np.abs(ux[1:-1, 1:-1]) + np.abs(uy[1:-1, 1:-1])).max())
    expect = np.float32(1) / (np.float32(b) + np.float32(3) * np.float32(np.float32(24) / np.float32(500)))
    assert e.check_cfl() == min(expect, np.float32(1))
    assert 0.3 < float(e.check_cfl()) < 0.8
    t0 = e.t
    e.set_xi(0.3, -0.3)
    for _ in range(3):
        dt = e.check_cfl()
        t_before = e.t
        e.update_adaptive()
        assert e.dt == dt and abs((e.t - t_before) - float(dt) / 24) < 1e-6
        assert np.isfinite(e.force()).all() and e.mg_iters()[0] >= 1
    assert e.t > t0
    # a fixed-dt step afterwards is NOT what the adaptive one did (dt really entered the step)
    f_adaptive = e.force()
    e2 = oracle.OracleEnv(literal=True)
    e2.set_state(init_state["ux"], init_state["uy"], init_state["p"])
    e2.set_xi(0.3, -0.3)
    for _ in range(3):
        e2.update2()
    assert e2.force()[0] != f_adaptive[0]
