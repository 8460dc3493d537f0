"""Multi-GPU plumbing: envs shard trivially, one process per GPU, no collective on the data path.

The only exchange is the per-iteration win/return statistics all-reduce (int64[16], 128 B) over NCCL
(NVLink 5 / NVSwitch) -- never inside an env-step.  The reference has no counterpart: it is single process and
keeps Python counters (game.py:24-26,140-141).
"""
import os

import torch
import torch.distributed as dist


def shard_range(rank, world_size, total_envs):
    """[lo, hi) of the global env ids owned by `rank`; contiguous, sizes differ by at most one."""
    if not (0 <= rank < world_size) or total_envs < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(total_envs, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_distributed(backend=None):
    """Join the torchrun world described by RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*; no-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


def allreduce_stats(stats, side_stream=None):
    """Sum the int64 stats vector over all ranks; returns a new tensor, `stats` itself is left per-rank.
    On CUDA the collective is issued on `side_stream` (if given) so it never serialises with env kernels."""
    out = stats.clone()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return out
    if out.is_cuda and side_stream is not None:
        side_stream.wait_stream(torch.cuda.current_stream(out.device))
        with torch.cuda.stream(side_stream):
            dist.all_reduce(out, op=dist.ReduceOp.SUM)
        out.record_stream(side_stream)
        torch.cuda.current_stream(out.device).wait_stream(side_stream)
    else:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out


def max_over_ranks(value, device):
    """max of a python float over ranks (device-timed step durations)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
