"""Multi-GPU plumbing: environments are independent, so a job shards the global env index range across ranks (one process per
GPU) with no data-path collective; the only collective is a tiny all-reduce of the episode statistics (NCCL on GPUs, gloo in the
CPU tests).  The device RNG is keyed by GLOBAL env index (pg_create env_id_offset), so results do not depend on the GPU count."""
from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[start, stop) of the global env indices owned by `rank`: contiguous, sizes differ by at most one."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, extra = divmod(total_envs, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_reduce_stats(stats: np.ndarray, device=None) -> np.ndarray:
    """Sum {episodes, successes, return_sum, length_sum} over all ranks (no-op without an initialised process group)."""
    t = torch.as_tensor(np.asarray(stats, dtype=np.float64), device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def make_sharded_env(task: str, total_envs: int, **kwargs):
    """PandaVecEnv for this rank's shard of a `total_envs`-env job (RANK / LOCAL_RANK / WORLD_SIZE from torchrun)."""
    import os
    from .vec_env import PandaVecEnv
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    start, stop = shard_range(total_envs, rank, world)
    return PandaVecEnv(task, stop - start, device=local, env_id_offset=start, **kwargs)
