"""world_size-2 gloo test of the N > 1 host path (env sharding, observation gather, max-over-ranks timing)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from rlfluidcontrol_b200.sharding import gather_observations, gather_observations_equal, max_over_ranks, shard_range
    import bench

    dist.init_process_group("gloo", rank=rank, world_size=world)
    e0, e1 = shard_range(16, rank, world)
    acts = bench.make_actions(3, e1 - e0, rank, e0=e0)
    # stand-in observation: a deterministic function of the global env id and this rank's first action
    obs = torch.tensor([[float(e), float(acts[0, e - e0, 0])] for e in range(e0, e1)], dtype=torch.float32)
    full = gather_observations(obs, world)
    # the preallocated equal-shard form bench.py uses inside its timed loop
    pre = torch.empty((world * obs.shape[0], 2), dtype=torch.float32)
    assert torch.equal(gather_observations_equal(obs, pre), full)
    # uneven shards (17 envs over 2 ranks: 8 + 9) keep the global env order
    u0, u1 = shard_range(17, rank, world)
    uneven = gather_observations(torch.arange(u0, u1, dtype=torch.float32).reshape(-1, 1), world)
    assert torch.equal(uneven[:, 0], torch.arange(17, dtype=torch.float32))
    t = max_over_ranks(1.0 + rank, world)
    out.put((rank, full.numpy().copy(), t, (e0, e1)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_and_timing():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, full0, t0, rng0), (r1, full1, t1, rng1) = res
    assert rng0 == (0, 8) and rng1 == (8, 16)
    assert np.array_equal(full0, full1) and full0.shape == (16, 2)
    assert np.array_equal(full0[:, 0], np.arange(16, dtype=np.float32))       # contiguous global env order
    assert t0 == t1 == 2.0                                                   # max over ranks
