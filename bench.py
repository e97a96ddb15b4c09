#!/usr/bin/env python
"""bench.py -- env-steps/s of batched AFCCylinder environments (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--envs-per-gpu B] [--impl b200|reference]

A "step" is one RL step of every environment in the batch: 16 solver steps (AFCCylinder.update2) with
the clientCFD.draw() force accumulation -> (Cl, Cd).  Workload (BASELINE configs[1]): default
AFCCylinder grid (384x192 cells), 256 environments per GPU, all resumed from init.bdim, synthetic
per-env action sequences.  N > 1 (torchrun, one rank per GPU): environments are independent, so each
rank owns its own 256 envs (weak scaling) and there is no data-path collective.

`value`  : device-resident throughput (actions already in HBM, rlfc_env_step_device, CUDA events).
`e2e`    : the same metric through the host-pointer C-ABI call rlfc_env_step (pinned host buffers,
           H2D actions + D2H obs/reward/done inside the timed region).
`roofline`: dominant kernel's algorithmic bytes / its CUDA-event duration vs measured HBM peak.
`cpu_baseline`: the oracle (literal C port of the reference step) on all host cores, bounded sample.
`--impl reference`: the CPU arm alone (reference Java cannot run here: no JVM; the literal C port of
           oracle/ stands in, one single-threaded process per host core).
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "env-steps/sec (batched BDIM cylinders)"
UNIT = "env-steps/s"


def workload_config(envs_per_gpu, n_gpus):
    return {
        "workload": "AFCCylinder default grid 384x192 (386x194 arrays), Re=500, "
                    f"{envs_per_gpu} batched envs per GPU from init.bdim, 16 solver steps per env-step, "
                    "synthetic per-env actions a=clip(0.8 sin(2 pi k/25 + 2 pi e/256)+0.1 N(0,1)) (env 0: config-1 sequence)",
        "envs_per_gpu": envs_per_gpu, "n_envs_total": envs_per_gpu * n_gpus, "substeps": 16,
        "mode": "exact (bit-identical to the oracle)", "l2": "per-GPU state >> 126 MB L2 (inputs larger than L2)",
        "parallelism": f"env-sharded x{n_gpus}, no data-path collective (observations all-gathered once per env-step)",
    }


def make_actions(n_steps, B, rank):
    """BASELINE config 2 synthetic actions."""
    rng = np.random.default_rng(1234 + rank)
    k = np.arange(n_steps)[:, None]
    e = (np.arange(B) + rank * B)[None, :]
    a1 = 0.8 * np.sin(2 * np.pi * k / 25 + 2 * np.pi * e / 256) + 0.1 * rng.standard_normal((n_steps, B))
    a2 = -0.8 * np.sin(2 * np.pi * k / 25 + 2 * np.pi * e / 256 + 1.0) + 0.1 * rng.standard_normal((n_steps, B))
    a = np.clip(np.stack([a1, a2], axis=-1), -1, 1).astype(np.float32)
    if rank == 0:
        kk = np.arange(n_steps)
        a[:, 0, 0] = (0.8 * np.sin(2 * np.pi * kk / 25.0)).astype(np.float32)
        a[:, 0, 1] = (-0.8 * np.sin(2 * np.pi * kk / 25.0 + 1.0)).astype(np.float32)
    return a


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle on all host cores (persistent single-threaded worker per core)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(conn, core, state_path):
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass
    from oracle import oracle_py as O
    st = O.read_bdimb(state_path)
    env = O.OracleEnv(literal=True)          # reference behaviour: coefficients rebuilt every step
    env.set_state(st["ux"], st["uy"], st["p"])
    k = 0
    conn.send("ready")
    while True:
        msg = conn.recv()
        if msg is None:
            break
        for _ in range(msg):
            a = (0.8 * np.sin(2 * np.pi * k / 25.0), -0.8 * np.sin(2 * np.pi * k / 25.0 + 1.0))
            env.env_step(a)
            k += 1
        conn.send(k)


class CpuArm:
    def __init__(self, cores=None):
        from oracle import oracle_py as O
        O.build()
        import rlfluidcontrol_b200 as R
        self.cores = cores or len(os.sched_getaffinity(0))
        ctx = mp.get_context("spawn")
        self.workers = []
        core_ids = sorted(os.sched_getaffinity(0))[: self.cores]
        for c in core_ids:
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_cpu_worker, args=(child, c, str(R.default_init_state())), daemon=True)
            p.start()
            self.workers.append((p, parent))
        for _, conn in self.workers:
            assert conn.recv() == "ready"

    def step(self, env_steps_each):
        """Every core advances its own environment by `env_steps_each` RL steps; returns wall seconds."""
        t0 = time.perf_counter()
        for _, conn in self.workers:
            conn.send(env_steps_each)
        for _, conn in self.workers:
            conn.recv()
        return time.perf_counter() - t0

    def close(self):
        for p, conn in self.workers:
            try:
                conn.send(None)
            except Exception:
                pass
        for p, _ in self.workers:
            p.join(timeout=5)


def cpu_baseline(sample_env_steps=4, reps=2):
    arm = CpuArm()
    arm.step(1)   # warm-up
    secs = [arm.step(sample_env_steps) for _ in range(reps)]
    arm.close()
    best = min(secs)
    return {
        "value": arm.cores * sample_env_steps / best, "unit": UNIT, "cores": arm.cores, "kind": "port",
        "sample": f"{arm.cores} single-threaded oracle processes (one per host core, literal C port of the Lilypad step, "
                  f"gcc -O2 -ffp-contract=off, coefficients rebuilt every step as the reference does), each advancing one "
                  f"AFCCylinder env from init.bdim by {sample_env_steps} env-steps (config-1 actions); best of {reps}",
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm()
    per_step = args.ref_env_steps
    for _ in range(args.warmup):
        arm.step(per_step)
    secs = [arm.step(per_step) for _ in range(args.steps)]
    arm.close()
    total = sum(secs)
    value = arm.cores * per_step * args.steps / total
    cfg = workload_config(args.envs_per_gpu, args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "port",
                         "sample": f"each step = {arm.cores} single-threaded processes (one per host core) x {per_step} env-steps of "
                                   "the literal C port of the reference step (no JVM in the image, so the Java itself cannot run)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks sampling
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` summary (profiles/r01_ncu_summary.json,
    same workload: 256 envs, default grid), or None."""
    p = ROOT / "profiles" / "r01_ncu_summary.json"
    if not p.exists():
        return None
    try:
        tab = json.loads(p.read_text())
    except Exception:
        return None
    for name, rec in tab.items():
        base = name.split("<")[0].replace("_rows", "").replace("_blk", "")
        if base == kernel or name.split("<")[0] == kernel:
            return float(rec["dram_traffic_B_per_launch"])
    return None


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from rlfluidcontrol_b200 import build as product_build
    if rank == 0:
        product_build.build()
    barrier()
    import rlfluidcontrol_b200 as R

    B, K, W = args.envs_per_gpu, args.steps, args.warmup
    stream = torch.cuda.Stream()
    env = R.AFCCylinderBatch(B, device=local, stream=stream.cuda_stream, init_time=-1.0)
    n_total = 2 * (W + K) + args.profile_steps
    acts_np = make_actions(n_total, B, rank)
    acts_dev = torch.from_numpy(acts_np).cuda()
    acts_host = torch.from_numpy(acts_np).pin_memory()
    obs_dev = torch.empty((B, 2), dtype=torch.float32, device="cuda")
    rew_dev = torch.empty(B, dtype=torch.float32, device="cuda")
    done_dev = torch.empty(B, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()

    obs_all = torch.empty((B * world, 2), dtype=torch.float32, device="cuda") if world > 1 else None

    def dev_step(k):
        env.step_device(acts_dev[k].data_ptr(), obs_dev.data_ptr(), rew_dev.data_ptr(), done_dev.data_ptr())
        if world > 1:   # the one exchange of the env-sharded path: every rank sees the whole batch's observations
            dist.all_gather_into_tensor(obs_all, obs_dev)

    step_idx = 0
    with torch.cuda.stream(stream):
        # ---- device-resident arm ----
        for _ in range(W):
            dev_step(step_idx); step_idx += 1
        barrier()
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        l0 = env.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(K):
            dev_step(step_idx); step_idx += 1
        e1.record(stream)
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        launches = env.launch_count - l0
        clk = clocks.stop() if rank == 0 else None
        obs_check = obs_dev.cpu().numpy()
        assert np.isfinite(obs_check).all(), "non-finite observation"

        # ---- end-to-end arm: host buffers through rlfc_env_step ----
        for _ in range(W):
            env.step(acts_host[step_idx].numpy()); step_idx += 1
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            obs, rew, done = env.step(acts_host[step_idx].numpy()); step_idx += 1
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        barrier()

        # ---- per-kernel CUDA-event pass (same workload, live) ----
        prof = []
        if rank == 0 and args.profile_steps > 0:
            env.set_profiling(True)
            for _ in range(args.profile_steps):   # rank 0 alone: no collective in here (the other ranks do not take part)
                env.step_device(acts_dev[step_idx].data_ptr(), obs_dev.data_ptr(), rew_dev.data_ptr(), done_dev.data_ptr())
                step_idx += 1
            torch.cuda.synchronize()
            prof = env.get_profile()
            env.set_profiling(False)
    mg = env.mg_iters()
    env.close()
    barrier()

    if rank == 0:
        total_envs = B * world
        value = total_envs * K / (ms / 1e3)
        e2e = total_envs * K / e2e_s
        peak, peak_src = measured_peak()
        kern = sorted(prof, key=lambda r: -r["ms"])
        tot_ms = sum(r["ms"] for r in kern) or 1.0
        roof = None
        table = []
        for r in kern:
            avg_ms = r["ms"] / max(r["launches"], 1)
            gbs = r["bytes_per_launch"] / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
            table.append({"kernel": r["name"], "share": round(r["ms"] / tot_ms, 4), "avg_ms": round(avg_ms, 4),
                          "launches_per_env_step": r["launches"] / max(args.profile_steps, 1),
                          "algorithmic_GBps": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 4)})
        if kern:
            top = kern[0]
            avg_ms = top["ms"] / max(top["launches"], 1)
            ach = top["bytes_per_launch"] / (avg_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": top["name"], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": ncu_traffic(top["name"]), "peak_source": peak_src, "share_of_step": top["ms"] / tot_ms,
                    "algorithmic_bytes_per_launch": top["bytes_per_launch"], "avg_launch_ms": avg_ms}
        # whole-step view against the SURVEY 8d model: 4 B * N_int * (28 + 13 (kP + kC)) per env per solver step
        nint = 384 * 192
        k_sum = float(mg.sum(axis=1).mean())
        model_bytes_step = 4.0 * nint * (28 + 13 * k_sum) * 16 * B
        whole = {"model_bytes_per_env_step_batch": model_bytes_step, "achieved_GBps": model_bytes_step / (ms / K * 1e-3) / 1e9}
        whole["frac_of_hbm_peak"] = whole["achieved_GBps"] / peak
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(B, world), "clocks": clk,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(B * 2 * 4), "d2h_bytes_per_step": int(B * (2 + 1 + 1) * 4)},
            "gpu_launches": int(launches), "roofline": roof, "whole_step_roofline": whole, "kernels": table,
            "mg_iters_per_solve": k_sum / 2,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.cpu_sample_env_steps)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--envs-per-gpu", type=int, default=256)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--profile-steps", type=int, default=1)
    ap.add_argument("--cpu-sample-env-steps", type=int, default=4)
    ap.add_argument("--ref-env-steps", type=int, default=2, help="env-steps per core per step in --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
