"""Multi-GPU sharding of batched environments: environments are independent, so global env ids are
split contiguously over ranks (one process per GPU) and the only exchange is the gather of the per-step
observations (2 floats per env) for a policy that wants the whole batch."""
from __future__ import annotations


def shard_range(n_envs_total: int, rank: int, world: int) -> tuple[int, int]:
    """[e0, e1) of the global env ids owned by `rank` (contiguous, sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    e0 = n_envs_total * rank // world
    e1 = n_envs_total * (rank + 1) // world
    return e0, e1


def gather_observations(obs, world: int, group=None):
    """All-gather the (B_local, 2) observation tensors of every rank into (B_total, 2), in global env order.
    Shards may differ in size by one (shard_range): they are padded to the largest shard for the collective and the
    padding is dropped again.  `out` (optional) is a preallocated (world * B_max, ...) buffer for the equal-shard case.
    Works with the gloo backend on CPU tensors and the nccl backend on CUDA tensors."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return obs
    sizes = torch.tensor([obs.shape[0]], dtype=torch.int64, device=obs.device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = [int(s.item()) for s in all_sizes]
    bmax = max(all_sizes)
    if obs.shape[0] < bmax:
        pad = torch.zeros((bmax - obs.shape[0],) + tuple(obs.shape[1:]), dtype=obs.dtype, device=obs.device)
        obs = torch.cat([obs, pad])
    out = torch.empty((world * bmax,) + tuple(obs.shape[1:]), dtype=obs.dtype, device=obs.device)
    dist.all_gather_into_tensor(out, obs.contiguous(), group=group)
    if min(all_sizes) == bmax:
        return out
    return torch.cat([out[r * bmax: r * bmax + all_sizes[r]] for r in range(world)])


def gather_observations_equal(obs, out, group=None):
    """The steady-state form used inside timed loops: equal shards, preallocated output, one collective."""
    import torch.distributed as dist

    dist.all_gather_into_tensor(out, obs, group=group)
    return out


def max_over_ranks(value: float, world: int, device="cpu") -> float:
    import torch
    import torch.distributed as dist

    if world == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
