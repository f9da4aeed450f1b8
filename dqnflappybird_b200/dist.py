"""Multi-GPU plumbing: one process per GPU, envs / replay shards partitioned, learner replicated.

SURVEY 8(e): env e of a job with ``total_envs`` lives on rank ``e * world // total_envs`` (contiguous blocks);
its frame ring and replay shard live on the same GPU; each rank samples ``batch // world`` transitions from its
own shard; Q-network, target network and Adam state are replicated; the only collective is one sum
all-reduce of the flat fp32 gradient vector per update (NCCL over NVLink on GPUs, gloo in the CPU tests).
Mean losses divide by the GLOBAL batch inside the loss kernel, so the reduced vector is the gradient of the
global loss and every rank applies the identical Adam step.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 or dist.is_initialized():
        return int(os.environ.get("RANK", "0")), world, int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group(backend)
    return dist.get_rank(), dist.get_world_size(), local_rank


def shard_envs(total_envs: int, rank: int, world: int):
    """(first_env_id, num_envs) of this rank: contiguous blocks, remainder spread over the first ranks."""
    base, rem = divmod(total_envs, world)
    n = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n


def owner_of_env(env: int, total_envs: int, world: int) -> int:
    base, rem = divmod(total_envs, world)
    cut = rem * (base + 1)
    return env // (base + 1) if env < cut else rem + (env - cut) // base


def local_batch(global_batch: int, world: int) -> int:
    if global_batch % world:
        raise ValueError("the minibatch must divide evenly over the ranks")
    return global_batch // world


def allreduce_gradients(flat_grads: torch.Tensor):
    """sum all-reduce of the flat gradient vector (898,722 floats = 3.6 MB; dueling 899,235)"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return flat_grads
