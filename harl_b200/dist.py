"""Thin helpers over torch.distributed for the rollout-sharded multi-GPU path.

One process per GPU; every rank owns N/G rollout threads and a full replica of all networks.
The only exchanges are sum-allreduces of small flat buffers (gradients, loss normalisers,
advantage / ValueNorm moments), so the replicas stay bit-identical after each optimiser step.
"""
import torch

# tests/dist_check_global_batch.py: run a runner with single-process semantics (no sharding, no exchanges) inside an
# initialised process group, to compare the sharded update with the same update on the whole global batch
FORCE_SINGLE = False


def is_dist():
    return not FORCE_SINGLE and torch.distributed.is_available() and torch.distributed.is_initialized()


def world_size():
    return torch.distributed.get_world_size() if is_dist() else 1


def rank():
    return torch.distributed.get_rank() if is_dist() else 0


def all_reduce_sum_(t):
    """In-place sum over ranks (no-op single-process)."""
    if is_dist() and world_size() > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
    return t


def shard_bounds(n_global, world, r):
    """Contiguous env shard [lo, hi) of rank r; n_global must divide evenly (weak scaling keeps it so)."""
    if n_global % world:
        raise ValueError(f"n_rollout_threads={n_global} is not divisible by world size {world}")
    per = n_global // world
    return r * per, (r + 1) * per
