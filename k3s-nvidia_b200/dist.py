"""One-process-per-GPU plumbing for the probe path (torch.distributed is plumbing, not product).

The HBM sweep and the GEMM probe are per-device and shard with NO data-path collective
("replicas only"); the only exchange step is the NVLink all-to-all, whose payload
moves through peer-mapped windows written by our own kernel (csrc/nvlink_a2a.cu).  What crosses
torch.distributed is control data only: 64-byte IPC handles, timings, verdicts.
Works on ``nccl`` (GPU box) and ``gloo`` (CPU tests, world_size 2).
"""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
import torch.distributed as dist


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend: str | None = None) -> tuple:
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def _dev():
    return torch.device("cuda", torch.cuda.current_device()) if dist.is_initialized() and dist.get_backend() == "nccl" else torch.device("cpu")


def barrier() -> None:
    if dist.is_initialized():
        dist.barrier()


_host_group = None


def host_barrier() -> None:
    """A barrier that keeps the GPUs idle while ranks wait.  dist.barrier() on NCCL parks an all-reduce kernel on every
    waiting rank's GPU; while rank 0 alone drives ALL GPUs through the single-process plugin entry (the step barrier of
    the exchange kernel needs every CTA of every device resident) nothing else may sit on their SMs.  Uses a gloo group
    over the same ranks, created on first use (collectively: every rank must call this the same number of times)."""
    global _host_group
    if not dist.is_initialized():
        return
    if dist.get_backend() != "nccl":
        dist.barrier()
        return
    if _host_group is None:
        _host_group = dist.new_group(backend="gloo")
    dist.barrier(group=_host_group)


def reduce_scalar(x: float, op: str = "max") -> float:
    if not dist.is_initialized():
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=_dev())
    dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN, "sum": dist.ReduceOp.SUM}[op])
    return float(t.item())


def all_gather_bytes(payload: bytes) -> List[bytes]:
    """Every rank contributes a fixed-size byte string (e.g. a 64-byte IPC handle)."""
    if not dist.is_initialized():
        return [payload]
    world = dist.get_world_size()
    mine = torch.tensor(list(payload), dtype=torch.uint8, device=_dev())
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return [bytes(t.cpu().tolist()) for t in out]


def all_gather_floats(vals: Sequence[float]) -> List[List[float]]:
    if not dist.is_initialized():
        return [list(vals)]
    world = dist.get_world_size()
    mine = torch.tensor(list(vals), dtype=torch.float64, device=_dev())
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return [t.cpu().tolist() for t in out]


def shard_units(n_units: int, rank: int, world: int) -> range:
    """Independent probe units (devices, sizes) -> contiguous shard for this rank (weak scaling
    uses one unit per rank; strong-scaling callers split a fixed list)."""
    per, rem = divmod(n_units, world)
    lo = rank * per + min(rank, rem)
    return range(lo, lo + per + (1 if rank < rem else 0))


def aggregate_bandwidth(bytes_per_rank: float, ms_per_rank: float) -> dict:
    """value = units all ranks processed / max-over-ranks time (the bench contract)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    t_max = reduce_scalar(ms_per_rank, "max")
    total = reduce_scalar(bytes_per_rank, "sum")
    return {"ms": t_max, "gbs": total / (t_max * 1e-3) / 1e9, "world": world}
