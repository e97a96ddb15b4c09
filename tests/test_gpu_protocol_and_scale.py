"""GPU tests beyond single-step parity: long traces, the full BASELINE batch, execution-mode equivalence
(CUDA graph vs eager, 1 vs several env groups), and the C client speaking the reference XML-RPC protocol."""
import os
import subprocess
import threading

import numpy as np
import pytest

from conftest import config1_actions

pytestmark = pytest.mark.gpu


def test_thousand_step_traces(rlfc, oracle, init_state):
    """north_star: drag / lift / sensor traces over 1000 solver steps within 1e-3 -- here exactly equal.  Config-1
    action sequence, one new action every 16 solver steps."""
    ref = oracle.OracleEnv(literal=False)
    ref.set_state(init_state["ux"], init_state["uy"], init_state["p"])
    with rlfc.AFCCylinderBatch(1) as env:
        for k in range(1000):
            a = config1_actions(k // 16)
            if k % 16 == 0:
                ref.set_xi(*a)
            f, pr = env.update2(a[None, :] if k % 16 == 0 else None, want_probes=True)
            ref.update2()
            fx, fy = ref.force()
            assert f[0, 0] == fx and f[0, 1] == fy, f"force differs at solver step {k}"
            if k % 50 == 0:
                assert np.array_equal(pr[0], ref.probes(32)), f"probes differ at step {k}"
        for nm, a, b in zip(("ux", "uy", "p"), env.get_fields(0), ref.get_state()):
            assert np.array_equal(a, b), nm


def test_full_batch_256_envs(rlfc, oracle, init_state):
    """BASELINE configs[1] size: 256 envs.  Envs fed identical actions stay identical to each other, and env 0 / env 255
    match lone oracle environments fed their own actions."""
    B = 256
    rng = np.random.default_rng(7)
    acts = np.clip(rng.normal(0, 0.5, (2, B, 2)), -1, 1).astype(np.float32)
    acts[:, 1::2, :] = acts[:, 0:1, :]            # every odd env mirrors env 0
    refs = {0: oracle.OracleEnv(literal=False), B - 2: oracle.OracleEnv(literal=False)}
    for r in refs.values():
        r.set_state(init_state["ux"], init_state["uy"], init_state["p"])
    with rlfc.AFCCylinderBatch(B, init_time=-1.0) as env:
        for k in range(2):
            obs, rew, done = env.step(acts[k])
            for e, r in refs.items():
                o = r.env_step(acts[k, e])
                assert obs[e, 0] == o[0] and obs[e, 1] == o[1], (k, e)
            assert np.array_equal(obs[1::2], np.repeat(obs[0:1], B // 2, axis=0))
        a = env.get_fields(0)
        b = env.get_fields(B - 1)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        assert (env.mg_iters() >= 1).all()


@pytest.mark.parametrize("mode", ["eager", "groups3"])
def test_execution_modes_are_equivalent(rlfc, mode, monkeypatch):
    """The eager path (one launch per kernel, host-side MG loop) and the multi-group CUDA-graph path (device-side
    WHILE loop) produce identical numbers, also from an impulsive start where MG needs several iterations."""
    B = 6
    acts = np.linspace(-1, 1, B * 2, dtype=np.float32).reshape(B, 2)

    def run(init):
        with rlfc.AFCCylinderBatch(B, init_state=init) as env:
            fs, its = [], []
            for k in range(3):
                fs.append(env.update2(acts if k == 0 else None).copy())
                its.append(env.mg_iters().copy())
            return fs, env.get_fields(B - 1), np.stack(its)

    base = {init: run(init) for init in ("default", None)}
    if mode == "eager":
        monkeypatch.setenv("RLFC_NO_GRAPH", "1")
    else:
        monkeypatch.setenv("RLFC_GROUPS", "3")
    for init in ("default", None):
        fs, fields, its = run(init)
        for a, b in zip(fs, base[init][0]):
            assert np.array_equal(a, b)
        for a, b in zip(fields, base[init][1]):
            assert np.array_equal(a, b)
        assert np.array_equal(its, base[init][2])
    assert base[None][2].max() > 1


def test_c_client_speaks_reference_protocol(rlfc, oracle, init_state, tmp_path):
    """rlfc_client (C, csrc/rlfc_client.c) against an XML-RPC peer: call order init, start_episode,
    request_stochastic_action("<Cl>_<Cd>") every 16 solver steps after t > initTime, train, save; payloads equal
    what clientCFD.pde would send (oracle-driven), trace file in the SaveScalar format."""
    from rlfluidcontrol_b200 import build_client
    from rlfluidcontrol_b200.peer.agent_server import AgentServer

    exe = build_client.build()
    srv = AgentServer("127.0.0.1", 0, agent="scripted", quiet=True, save_dir=str(tmp_path / "save"), data_dir=str(tmp_path / "data"))
    th = threading.Thread(target=srv.serve_forever, daemon=True)
    th.start()
    n_rl = 3
    init_time = 0.1                      # 14 warm-up solver steps (t = 0.0075 k > 0.1 from k = 14)
    t_end = np.float32(0)
    steps = 0
    while True:                          # episode length so that exactly n_rl actions are requested
        steps += 1
        t_end = np.float32(t_end + np.float32(np.float32(np.float32(0.0075) * 24) / 24))
        if steps == 13 + 16 * n_rl:
            break
    try:
        res = subprocess.run([str(exe), "--host", "127.0.0.1", "--port", str(srv.port), "--episodes", "1", "--time", repr(float(t_end)),
                              "--init-time", str(init_time), "--init", str(rlfc.default_init_state()), "--save-dir", str(tmp_path / "saved"),
                              "--train-steps", "2", "--quiet"], capture_output=True, text=True, timeout=300)
    finally:
        srv.server.shutdown()
    assert res.returncode == 0, res.stderr
    names = [c[0] for c in srv.calls]
    assert names == ["init", "start_episode"] + ["request_stochastic_action"] * n_rl + ["train", "save"], names
    assert srv.calls[0][1] == -1 and srv.calls[1][1] == -1 and srv.calls[-2][1] == 2
    # expected payloads from the oracle driven exactly like clientCFD.draw()
    ref = oracle.OracleEnv(literal=False)
    ref.set_state(init_state["ux"], init_state["uy"], init_state["p"])
    L = rlfc.load_library()
    import ctypes as C
    L.rlfc_format_float_java.argtypes = [C.c_float, C.c_char_p, C.c_int]
    buf = C.create_string_buffer(64)

    def jf(v):
        L.rlfc_format_float_java(np.float32(v), buf, 64)
        return buf.value.decode()

    expected, k = [], 0
    for _ in range(steps):
        o = ref.driver_step(init_time)
        if o is not None:
            expected.append(f"{jf(o[0])}_{jf(o[1])}")
            ref.set_xi(*config1_actions(k))          # the scripted peer's k-th reply, parsed like Float.parseFloat
            k += 1
    sent = [c[1] for c in srv.calls if c[0] == "request_stochastic_action"]
    assert sent == expected
    lines = (tmp_path / "saved" / "1.txt").read_text().splitlines()
    assert lines[0].startswith("%% Force and pressure") and len(lines) == 4 + steps
    assert len(lines[4].split()) == 5 + 32


def test_env_step_default_init_time_matches_reference_cadence(rlfc, oracle, init_state):
    """rlfc_env_step with the DEFAULT init_time = 1 (clientCFD.pde:5): the first call after a reset runs uncontrolled
    until the reference's first observation (133 steps with t <= 1, then one 16-step window; the action passed is
    ignored, as the sketch has not asked for one), every later call is 16 solver steps under one action.  Compared with
    the oracle's draw() loop (ora_driver_step) driven the way clientCFD.pde:35-55 drives it."""
    ref = oracle.OracleEnv(literal=False)
    ref.set_state(init_state["ux"], init_state["uy"], init_state["p"])

    def ref_until_obs(limit=400):
        for k in range(1, limit + 1):
            o = ref.driver_step(1.0)
            if o is not None:
                return o, k
        raise AssertionError("no observation")

    with rlfc.AFCCylinderBatch(2) as env:                 # default config: init_time = 1
        obs, _, done = env.step(np.array([[0.7, -0.7], [0.0, 0.0]], np.float32))   # env 0's action must be ignored
        o, k = ref_until_obs()
        assert k == 149                                   # 133 steps with t <= 1, then callLearn 16 -> 0
        assert obs[0, 0] == o[0] and obs[0, 1] == o[1] and obs[1, 0] == o[0] and obs[1, 1] == o[1]
        assert env.running() == 0 and done.tolist() == [0, 0]
        for step in range(3):
            a = config1_actions(step)
            obs, _, _ = env.step(np.stack([a, a]))
            ref.set_xi(a[0], a[1])                        # callAction's answer, applied from the next frame on
            o, k = ref_until_obs()
            assert k == 16
            assert obs[0, 0] == o[0] and obs[0, 1] == o[1], (step, obs[0], o)
        # env 1 is reset alone: it catches up over several rounds (uncontrolled) while env 0 stays put
        ux0 = env.get_fields(0)[0].copy()
        t0 = env.t.copy()
        env.reset([1], reset_accumulators=True)
        obs, _, _ = env.step(np.zeros((2, 2), np.float32))
        t1 = env.t
        # env 0 advanced exactly 16 solver steps, env 1 149 from t = 0
        assert abs(t1[0] - (t0[0] + 16 * 0.0075)) < 1e-4 and abs(t1[1] - 149 * 0.0075) < 1e-4
        ref2 = oracle.OracleEnv(literal=False)
        ref2.set_state(init_state["ux"], init_state["uy"], init_state["p"])
        for k in range(149):
            o2 = ref2.driver_step(1.0)
        assert obs[1, 0] == o2[0] and obs[1, 1] == o2[1]
        assert np.array_equal(env.get_fields(1)[0], ref2.get_state()[0])
        assert not np.array_equal(env.get_fields(0)[0], ux0)


def test_episode_end_freezes_and_flags(rlfc):
    """An environment past episode_time takes no more steps (clientCFD.pde:36); a diverged one raises its flag."""
    with rlfc.AFCCylinderBatch(2, init_time=-1.0, episode_time=0.2) as env:   # 0.2 / 0.0075 = 26.7 solver steps
        z = np.zeros((2, 2), np.float32)
        _, _, d = env.step(z)
        assert d.tolist() == [0, 0]
        _, _, d = env.step(z)                              # t reaches 0.2025 at step 27: frozen there, mid-window
        assert d.tolist() == [1, 1]
        t = env.t.copy()
        assert abs(t[0] - 27 * 0.0075) < 1e-5
        o1, _, d = env.step(z)
        assert np.array_equal(env.t, t) and d.tolist() == [1, 1]
        assert env.flags().tolist() == [0, 0]
        env.reset([0])
        ux, uy, p = env.get_fields(0)
        p[10, 10] = np.nan
        env.set_fields(0, p=p)
        env.step(z)
        assert env.flags().tolist() == [1, 0]
