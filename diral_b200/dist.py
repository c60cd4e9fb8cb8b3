"""Multi-GPU plumbing: the env batch shards by contiguous env ranges, one process per GPU, and the
only exchange is one all-reduce(sum) of the 110-element end-of-episode metric vector
(reference accounting at main_test.py:178-182,226-231)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world: int):
    """Contiguous, balanced env range [lo, hi) of `rank`; the first (total % world) ranks get one more."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    base, extra = divmod(int(total_envs), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_from_env(backend: str | None = None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*); returns (rank, world, local_rank)."""
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def all_reduce_metrics(vec: torch.Tensor, async_op: bool = False):
    """Sum the per-device metric vector over all ranks in place (NCCL over NVLink on GPUs, gloo on
    CPU tensors).  ~1 KB: pure latency, so callers issue it on a side stream at episode end."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        return dist.all_reduce(vec, op=dist.ReduceOp.SUM, async_op=async_op)
    return None


def metrics_dict(vec: torch.Tensor):
    from .env import METRIC_FIELDS
    v = vec.detach().cpu().tolist()
    d = {k: v[i] for i, k in enumerate(METRIC_FIELDS[:6])}
    d["env_slots"] = v[6]
    d["prr"] = v[2] / v[3] if v[3] else float("nan")
    d["information_age"] = v[10:110]
    return d
