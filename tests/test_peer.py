"""CPU tests of the protocol-compatible agent peer (rlfluidcontrol_b200/peer/agent_server.py): the TD3 learner
actually learns from what the RPC methods record, checkpoints round-trip, and batched environments keep one record stream
each (server/server.py:157-165 pairs CONSECUTIVE records as transitions, so interleaving environments would corrupt them)."""
import threading
import xmlrpc.client

import numpy as np
import pytest

from rlfluidcontrol_b200.peer.agent_server import AgentServer, TD3Agent, reward_func


def serve(tmp_path, **kw):
    srv = AgentServer("127.0.0.1", 0, quiet=True, save_dir=str(tmp_path / "save"), data_dir=str(tmp_path / "data"), **kw)
    th = threading.Thread(target=srv.serve_forever, daemon=True)
    th.start()
    return srv, xmlrpc.client.ServerProxy(f"http://127.0.0.1:{srv.port}", allow_none=True)


def test_td3_hyperparameters_match_the_reference():
    """server/agent_TD3.py:12-27"""
    a = TD3Agent(seed=0)
    assert (a.lr, a.gamma, a.tau, a.bs, a.bfs, a.d) == (1e-4, 0.99, 0.005, 512, 1_000_000, 2)
    assert (a.explore_noise_size, a.smooth_noise, a.smooth_clip) == (0.1, 0.2, 0.5)
    sizes = [m.out_features for m in a.actor if hasattr(m, "out_features")]
    assert sizes == [256, 256, 2]


def test_reward_is_the_reference_formula():
    """server/server.py:61-65: -Cd_next - pi/8 * 0.0097 * 3.66^3 * sum |a|^3"""
    s = np.array([[0.3, 1.2]])
    a = np.array([[0.5, -1.0]])
    assert abs(reward_func(s, a) - (-1.2 - np.pi / 8 * 0.0097 * 3.66 ** 3 * (0.125 + 1.0))) < 1e-12


def test_batch_records_are_paired_per_environment(tmp_path):
    """request_batch_action keeps one stream per environment; train() turns consecutive records of ONE environment into
    transitions (never env e's state with env e+1's next state)."""
    srv, cl = serve(tmp_path, agent="scripted")
    try:
        assert cl.init(-1) is False and cl.start_episode(-1) is True
        B, T = 3, 5
        for k in range(T):
            payload = ";".join(f"{10 * e + k}.0_{100 * e + k}.5" for e in range(B))
            reply = cl.request_batch_action(payload)
            assert len(reply.split(";")) == B
        assert cl.train(0) is True
    finally:
        srv.server.shutdown()
    buf = srv.agent.replay_buffer
    assert buf.size == B * (T - 1)
    for i in range(buf.size):
        e, k = divmod(i, T - 1)
        assert buf.obs1[i].tolist() == [10 * e + k, 100 * e + k + 0.5]
        assert buf.obs2[i].tolist() == [10 * e + k + 1, 100 * e + k + 1.5]          # same environment, next record
        assert buf.done[i, 0] == 0
    # an empty field = an environment without an observation this round: action 0, nothing recorded


def test_finish_envs_and_empty_fields(tmp_path):
    srv, cl = serve(tmp_path, agent="scripted")
    try:
        cl.start_episode(-1)
        cl.request_batch_action("1.0_2.0;3.0_4.0")
        cl.request_batch_action("1.5_2.5;3.5_4.5")
        cl.finish_envs("1")                          # env 1's episode is over: one transition stored, stream cleared
        assert srv.agent.replay_buffer.size == 1 and srv.agent.replay_buffer.obs1[0].tolist() == [3.0, 4.0]
        reply = cl.request_batch_action("1.75_2.75;")   # env 1 in its uncontrolled start
        assert reply.split(";")[1] == "0.0_0.0"
        assert 1 not in srv.batch_records and len(srv.batch_records[0]) == 3
    finally:
        srv.server.shutdown()


def test_td3_trains_and_checkpoints(tmp_path):
    """Replay buffer filled through the RPC surface (16 environments x 40 RL steps of synthetic observations), train(50)
    changes the actor, save -> restore round-trips networks and buffer (server.py:148-198)."""
    import torch
    srv, cl = serve(tmp_path, agent="td3", seed=3)
    rng = np.random.default_rng(0)
    try:
        cl.init(-1)
        cl.start_episode(-1)
        B, T = 16, 40
        for k in range(T):
            obs = np.stack([0.1 * rng.standard_normal(B), 1.1 + 0.1 * rng.standard_normal(B)], axis=1)
            reply = cl.request_batch_action(";".join(f"{o[0]!r}_{o[1]!r}" for o in obs.tolist()))
            acts = np.array([[float(v) for v in p.split("_")] for p in reply.split(";")])
            assert acts.shape == (B, 2) and np.abs(acts).max() <= 1.0
        before = [p.detach().clone() for p in srv.agent.actor.parameters()]
        q_before = [p.detach().clone() for p in srv.agent.q1.parameters()]
        assert cl.train(50) is True
        assert srv.agent.replay_buffer.size == B * (T - 1) >= srv.agent.bs
        assert srv.agent.train_count == 50
        assert any(not torch.equal(a, b) for a, b in zip(before, srv.agent.actor.parameters()))
        assert any(not torch.equal(a, b) for a, b in zip(q_before, srv.agent.q1.parameters()))
        assert cl.save(50) is True
        ep = srv.agent.episode_count
        trained = [p.detach().clone() for p in srv.agent.actor.parameters()]
        probe = np.array([[0.05, 1.15]])
        a_trained = srv.agent.get_action(probe, stochastic=False)
        srv.agent.reset_agent()                       # fresh networks, empty buffer
        assert srv.agent.replay_buffer.size == 0
        assert cl.restore(ep) is True
        assert all(torch.equal(a, b) for a, b in zip(trained, srv.agent.actor.parameters()))
        assert srv.agent.replay_buffer.size == B * (T - 1)
        assert np.array_equal(srv.agent.get_action(probe, stochastic=False), a_trained)
        # the deterministic single-environment method of the unchanged protocol answers from the same actor
        r = cl.request_deterministic_action("0.05_1.15")
        assert np.allclose([float(v) for v in r.split("_")], a_trained[0], atol=1e-6)
    finally:
        srv.server.shutdown()
