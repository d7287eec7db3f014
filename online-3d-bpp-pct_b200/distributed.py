"""Multi-GPU sharding of the environment batch: one process per GPU, contiguous global env ranges, NO collective in
the step.  Environments are independent (the reference runs them as isolated worker processes,
wrapper/shmem_vec_env.py:48-57); every per-env random stream is keyed by the GLOBAL env index, so the trajectories of
a job do not depend on how many GPUs it is split over.

`gather_observations` is the single optional collective (NCCL all-gather of the observation buffer) for a learner
that wants the whole rollout batch on every rank.
"""
import torch


def shard_range(n_total, world_size, rank):
    """Global env indices [base, base + count) owned by `rank`."""
    base = n_total * rank // world_size
    return base, n_total * (rank + 1) // world_size - base


def make_sharded_vec_env(n_total, setting, rank=None, world_size=None, **kw):
    """PctVecEnv over this rank's slice of `n_total` global envs (device = local rank unless given)."""
    import torch.distributed as dist
    from .vec_env import PctVecEnv
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    base, count = shard_range(n_total, world_size, rank)
    kw.setdefault("device", rank % max(torch.cuda.device_count(), 1))
    return PctVecEnv(count, setting, env_id_base=base, **kw)


def gather_observations(obs, world_size=None, group=None):
    """all-gather of equally sized per-rank observation buffers -> (world * n_local, obs_len), rank-major
    (== global env order for equal shards).  NCCL on GPUs, works on gloo/CPU tensors for tests."""
    import torch.distributed as dist
    world_size = world_size or dist.get_world_size(group)
    out = torch.empty((world_size * obs.shape[0],) + tuple(obs.shape[1:]), dtype=obs.dtype, device=obs.device)
    if hasattr(dist, "all_gather_into_tensor") and obs.is_cuda:
        dist.all_gather_into_tensor(out, obs.contiguous(), group=group)
    else:
        parts = list(out.chunk(world_size, dim=0))
        dist.all_gather(parts, obs.contiguous(), group=group)
    return out
