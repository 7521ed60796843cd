"""Multi-GPU plumbing: one process per GPU, env replicas sharded across ranks (each rank
owns `num_envs` independent replicas, RNG seed = seed + rank as in trainer_base.py:249-251),
NO collective on the rollout path, and ONE flat all-reduce (mean) of the gradients of all
trained policies per training iteration.

The reference spawns one process per GPU too but synchronises gradients with DDP over a
*gloo* group through host memory (training/utils/process_group_torch.py:6-20,
trainer_a2c.py:137-146); here the group is NCCL over NVLink/NVSwitch and the ~0.8 MB of
gradients travel as a single latency-bound message.  The same code runs on a gloo group
for the CPU unit tests.
"""
import os

import torch
import torch.distributed as dist


def init_process_group(backend=None, device_id=None):
    """Initialise from the torchrun environment (RANK / WORLD_SIZE / MASTER_*)."""
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kwargs = {}
    if backend == "nccl" and device_id is not None:
        kwargs["device_id"] = torch.device("cuda", device_id)
    dist.init_process_group(backend, **kwargs)
    return dist.get_rank(), dist.get_world_size()


def flat_allreduce_mean_(tensors, world_size, flat_buffer=None):
    """In-place mean of `tensors` across ranks with a single all-reduce.  Returns the flat
    buffer so callers can reuse it across iterations."""
    tensors = [t for t in tensors if t is not None]
    if not tensors or world_size <= 1:
        return flat_buffer
    n = sum(t.numel() for t in tensors)
    if flat_buffer is None or flat_buffer.numel() != n or flat_buffer.device != tensors[0].device:
        flat_buffer = torch.empty(n, device=tensors[0].device, dtype=tensors[0].dtype)
    torch.cat([t.reshape(-1) for t in tensors], out=flat_buffer)
    dist.all_reduce(flat_buffer, op=dist.ReduceOp.SUM)
    flat_buffer.div_(world_size)
    off = 0
    for t in tensors:
        t.copy_(flat_buffer[off:off + t.numel()].view_as(t))
        off += t.numel()
    return flat_buffer


def shard_of(total_envs, rank, world_size):
    """Contiguous block of env replicas owned by `rank` (strong-scaling split of a global
    env count; weak scaling simply gives every rank `num_envs`)."""
    base, rem = divmod(total_envs, world_size)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def max_over_ranks(value, device=None):
    """Timing helper: the slowest rank defines the step time."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64,
                     device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
