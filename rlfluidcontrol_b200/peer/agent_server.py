"""XML-RPC agent peer speaking the reference protocol (server/server.py:51-59, CFD mode), on PyTorch.

The reference server cannot be imported in this image (TensorFlow 1 is absent and `np.asfarray`, which
server.py:118 calls, no longer exists in numpy 2), so this module provides the same eight RPC methods with
the same payload formats and the same TD3 hyper-parameters (agent_TD3.py:12-27: two 256-unit ReLU layers,
tanh actor, twin critics, target-policy smoothing 0.2 clipped at 0.5, policy delay 2, Adam 1e-4, batch 512,
tau 0.005, gamma 0.99, Gaussian exploration 0.1) and the same reward (server.py:61-65).  It is what the
protocol tests drive rlfc_client against; the unchanged reference server is a drop-in replacement for it
wherever TensorFlow 1 is available.

Extension (not in the reference): `request_batch_action("Cl_Cd;Cl_Cd;...")` answers for many environments in one
call and keeps one record stream per environment, so batched environments neither serialise through one RPC
each nor corrupt the consecutive-record transition pairing of server.py:157-165.
"""
from __future__ import annotations

import argparse
import os
import pickle
from datetime import datetime, timezone
from xmlrpc.server import SimpleXMLRPCServer

import numpy as np


def reward_func(next_state, action):
    """server.py:61-65"""
    return -next_state[0, 1] - np.pi * (1 / 8) * 0.0097 * (3.66 ** 3) * np.sum(np.abs(action) ** 3)


class ReplayBuffer:
    """FIFO replay buffer with the reference layout (utils.py:38-66)."""

    def __init__(self, obs_dim, act_dim, size):
        self.obs1 = np.zeros((size, obs_dim), np.float32)
        self.obs2 = np.zeros((size, obs_dim), np.float32)
        self.acts = np.zeros((size, act_dim), np.float32)
        self.rews = np.zeros((size, 1), np.float32)
        self.done = np.zeros((size, 1), np.float32)
        self.ptr = self.size = 0
        self.max_size = size

    def store(self, obs, act, rew, next_obs, done):
        k = self.ptr
        self.obs1[k], self.obs2[k], self.acts[k], self.rews[k], self.done[k] = obs, next_obs, act, rew, done
        self.ptr = (k + 1) % self.max_size
        self.size = min(self.size + 1, self.max_size)

    def sample(self, n):
        idx = np.random.randint(0, self.size, size=n)
        return self.obs1[idx], self.acts[idx], self.rews[idx], self.done[idx], self.obs2[idx]


class TD3Agent:
    """TD3 with the reference's hyper-parameters (agent_TD3.py:11-27, 95-155)."""

    def __init__(self, state_dim=2, action_dim=2, device="cpu", seed=None):
        import torch
        import torch.nn as nn
        self.torch = torch
        if seed is not None:
            torch.manual_seed(seed)
            np.random.seed(seed)
        self.lr, self.gamma, self.tau, self.bs, self.bfs, self.d = 1e-4, 0.99, 0.005, 512, 1_000_000, 2
        self.explore_noise_size, self.smooth_noise, self.smooth_clip = 0.1, 0.2, 0.5
        self.state_dim, self.action_dim, self.device = state_dim, action_dim, device

        def mlp(i, o, last):
            return nn.Sequential(nn.Linear(i, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, o), last)

        self.actor = mlp(state_dim, action_dim, nn.Tanh()).to(device)
        self.q1 = mlp(state_dim + action_dim, 1, nn.Identity()).to(device)
        self.q2 = mlp(state_dim + action_dim, 1, nn.Identity()).to(device)
        import copy
        self.actor_t, self.q1_t, self.q2_t = (copy.deepcopy(m) for m in (self.actor, self.q1, self.q2))
        self.opt_q = torch.optim.Adam(list(self.q1.parameters()) + list(self.q2.parameters()), lr=self.lr)
        self.opt_p = torch.optim.Adam(self.actor.parameters(), lr=self.lr)
        self.reset_agent(reinit=False)

    def reset_agent(self, reinit=True):
        if reinit:
            for m in (self.actor, self.q1, self.q2):
                for layer in m:
                    if hasattr(layer, "reset_parameters"):
                        layer.reset_parameters()
            for t, s in ((self.actor_t, self.actor), (self.q1_t, self.q1), (self.q2_t, self.q2)):
                t.load_state_dict(s.state_dict())
        self.replay_buffer = ReplayBuffer(self.state_dim, self.action_dim, self.bfs)
        self.step_count = self.total_step_count = self.train_count = self.episode_count = 0

    def reset_episode(self):
        self.step_count = self.train_count = 0
        self.episode_count += 1

    def get_action(self, state, stochastic=True):
        torch = self.torch
        with torch.no_grad():
            a = self.actor(torch.as_tensor(state, dtype=torch.float32, device=self.device)).cpu().numpy()
        if stochastic:
            a = np.clip(a + np.random.normal(0, self.explore_noise_size, a.shape), -1, 1)
        return a

    def train_iter(self):
        torch = self.torch
        if self.bs > self.replay_buffer.size:
            return
        o1, a, r, d, o2 = (torch.as_tensor(x, device=self.device) for x in self.replay_buffer.sample(self.bs))
        with torch.no_grad():
            noise = (torch.randn_like(a) * self.smooth_noise).clamp(-self.smooth_clip, self.smooth_clip)
            a2 = (self.actor_t(o2) + noise).clamp(-1, 1)
            y = r + self.gamma * (1 - d) * torch.minimum(self.q1_t(torch.cat([o2, a2], 1)), self.q2_t(torch.cat([o2, a2], 1)))
        oa = torch.cat([o1, a], 1)
        q_loss = ((self.q1(oa) - y) ** 2).mean() + ((self.q2(oa) - y) ** 2).mean()
        self.opt_q.zero_grad(); q_loss.backward(); self.opt_q.step()
        if self.total_step_count % self.d == 0:
            p_loss = -self.q1(torch.cat([o1, self.actor(o1)], 1)).mean()
            self.opt_p.zero_grad(); p_loss.backward(); self.opt_p.step()
            with torch.no_grad():
                for t, s in ((self.actor_t, self.actor), (self.q1_t, self.q1), (self.q2_t, self.q2)):
                    for pt, ps in zip(t.parameters(), s.parameters()):
                        pt.mul_(1 - self.tau).add_(self.tau * ps)
        self.train_count += 1
        self.total_step_count += 1


class ScriptedAgent:
    """Deterministic stand-in policy for protocol tests: a_k = (0.8 sin(2 pi k/25), -0.8 sin(2 pi k/25 + 1))."""

    def __init__(self, *a, **k):
        self.k = 0
        self.episode_count = 0
        self.replay_buffer = ReplayBuffer(2, 2, 1024)

    def reset_agent(self): self.k = 0
    def reset_episode(self): self.episode_count += 1

    def get_action(self, state, stochastic=True):
        a = np.array([[0.8 * np.sin(2 * np.pi * self.k / 25.0), -0.8 * np.sin(2 * np.pi * self.k / 25.0 + 1.0)]], np.float32)
        self.k += 1
        return np.repeat(a, len(state), axis=0)

    def train_iter(self): pass


class AgentServer:
    """Same RPC surface as server/server.py `Server` in "CFD" mode with the NoneFilter."""

    def __init__(self, host="localhost", port=8000, agent="td3", save_dir="save", data_dir="save_data", eval_dir="save_eval",
                 seed=None, quiet=False):
        self.agent = TD3Agent(seed=seed) if agent == "td3" else ScriptedAgent()
        self.state_dim, self.action_dim = 2, 2
        self.quiet = quiet
        self.save_model_dir, self.save_data_dir, self.save_eval_dir = save_dir, data_dir, eval_dir
        self.state_record, self.unfiltered_state_record, self.action_record = [], [], []
        self.batch_records = {}
        self.calls = []                                   # (method, payload) log used by the tests
        self.server = SimpleXMLRPCServer((host, port), logRequests=False, allow_none=True)
        self.port = self.server.server_address[1]
        for name in ("init", "start_episode", "request_stochastic_action", "request_deterministic_action", "train", "save",
                     "save_eval", "restore", "request_batch_action", "finish_envs"):
            self.server.register_function(getattr(self, "_" + name), name)

    def _stamp(self, s):
        if not self.quiet:
            print("UTC " + datetime.now(timezone.utc).isoformat(sep=" ", timespec="milliseconds") + " " + s, flush=True)

    # ---- server.py:92-112 ----
    def _init(self, episode_count):
        self.calls.append(("init", episode_count))
        try:
            self._restore(episode_count)
            return True
        except Exception:
            self.agent.reset_agent()
            self._stamp("Initialized!")
            return False

    def _start_episode(self, raw_data):
        self.calls.append(("start_episode", raw_data))
        self.state_record, self.unfiltered_state_record, self.action_record = [], [], []
        self.batch_records = {}
        self.agent.reset_episode()
        self._stamp("Epsode Start!")
        return True

    # ---- server.py:114-146 ----
    def _request_action(self, raw_data, stochastic):
        state = np.asarray(raw_data.split("_"), dtype=float)[None, :]
        action = self.agent.get_action(state, stochastic=stochastic)
        self.unfiltered_state_record.append(state)
        self.state_record.append(state)
        self.action_record.append(action)
        return "_".join(str(i) for i in action[0, :])

    def _request_stochastic_action(self, raw_data):
        self.calls.append(("request_stochastic_action", raw_data))
        return self._request_action(raw_data, True)

    def _request_deterministic_action(self, raw_data):
        self.calls.append(("request_deterministic_action", raw_data))
        return self._request_action(raw_data, False)

    def _request_batch_action(self, raw_data):
        """Extension: "Cl_Cd;Cl_Cd;..." -> "a1_a2;a1_a2;..." with one record stream per environment.  An empty field
        ("Cl_Cd;;Cl_Cd") marks an environment that has no observation this round (still in its uncontrolled start): it
        gets the action "0.0_0.0" and nothing is recorded for it."""
        self.calls.append(("request_batch_action", raw_data))
        fields = raw_data.split(";")
        live = [e for e, f in enumerate(fields) if f]
        out = ["0.0_0.0"] * len(fields)
        if live:
            states = np.array([[float(v) for v in fields[e].split("_")] for e in live])
            actions = self.agent.get_action(states, stochastic=True)
            for e, s, a in zip(live, states, actions):
                self.batch_records.setdefault(e, []).append((s[None, :], a[None, :]))
                out[e] = "_".join(str(i) for i in a)
        return ";".join(out)

    def _finish_envs(self, raw_data):
        """Extension: "3;7" -- the episodes of these environments are over (batched auto-reset): their record streams
        become transitions now (consecutive records of ONE environment, server.py:157-165) and start afresh."""
        self.calls.append(("finish_envs", raw_data))
        for e in (int(v) for v in raw_data.split(";") if v):
            rec = self.batch_records.pop(e, [])
            self._store_transitions([s for s, _ in rec], [a for _, a in rec])
        return True

    # ---- server.py:148-203 ----
    def _store_transitions(self, states, actions):
        for i in range(len(states) - 1):
            r = reward_func(states[i + 1], actions[i])
            self.agent.replay_buffer.store(states[i], actions[i][0, :self.action_dim], r, states[i + 1], 0)

    def _train(self, steps):
        self.calls.append(("train", steps))
        os.makedirs(self.save_data_dir, exist_ok=True)
        np.savez(os.path.join(self.save_data_dir, f"data_{self.agent.episode_count}.npz"), state=np.array(self.state_record),
                 unfiltered_state=np.array(self.unfiltered_state_record), action=np.array(self.action_record))
        self._stamp("Length of Record: " + str(len(self.state_record)))
        self._store_transitions(self.state_record, self.action_record)
        for rec in self.batch_records.values():
            self._store_transitions([s for s, _ in rec], [a for _, a in rec])
        self.batch_records = {}
        self._stamp("Training Start!")
        for _ in range(steps):
            self.agent.train_iter()
        self._stamp("Training End!")
        return True

    def _save_eval(self, episode_count, rep_count):
        os.makedirs(self.save_eval_dir, exist_ok=True)
        np.savez(os.path.join(self.save_eval_dir, f"data_{episode_count}_{rep_count}.npz"), state=np.array(self.state_record),
                 unfiltered_state=np.array(self.unfiltered_state_record), action=np.array(self.action_record))
        return True

    def _save(self, dummy=None):
        self.calls.append(("save", dummy))
        os.makedirs(self.save_model_dir, exist_ok=True)
        with open(os.path.join(self.save_model_dir, f"{self.agent.episode_count}.pickle"), "wb") as f:
            state = {"buffer": self.agent.replay_buffer, "episode": self.agent.episode_count}
            if isinstance(self.agent, TD3Agent):
                state["nets"] = {k: getattr(self.agent, k).state_dict() for k in ("actor", "q1", "q2", "actor_t", "q1_t", "q2_t")}
            pickle.dump(state, f)
        self._stamp(f"Saved Episode {self.agent.episode_count}!")
        return True

    def _restore(self, episode_count):
        with open(os.path.join(self.save_model_dir, f"{episode_count}.pickle"), "rb") as f:
            state = pickle.load(f)
        self.agent.replay_buffer = state["buffer"]
        self.agent.episode_count = episode_count
        for k, sd in state.get("nets", {}).items():
            getattr(self.agent, k).load_state_dict(sd)
        self._stamp(f"Restored from Episode {episode_count}!")
        return True

    def serve_forever(self):
        self._stamp("Server Listening...")
        self.server.serve_forever()


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="agent peer for the CFD client (reference protocol)")
    ap.add_argument("--host", default="localhost")
    ap.add_argument("--port", type=int, default=8000)
    ap.add_argument("--agent", choices=["td3", "scripted"], default="td3")
    args = ap.parse_args()
    AgentServer(args.host, args.port, args.agent).serve_forever()
