"""Distributed wrapper for the sampling path (reference: rqvae/utils/dist.py:18-103).

One process per GPU, `torch.distributed` (NCCL on GPUs, gloo in CPU tests), env:// rendezvous.  Only what the sampling
scripts use: ``initialize``, ``dataparallel_and_sync`` (returns an object with ``.module``; weights are made identical by a
single flat broadcast per dtype instead of the reference's ~780 per-tensor broadcasts), ``all_gather_cat`` and
``shard_batch`` (the label-grid split of main_sampling_fid.py:196-206)."""
import datetime
import os
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass
class DistEnv:
    world_size: int
    world_rank: int
    local_rank: int
    num_gpus: int
    master: bool
    device_name: str


def initialize(args=None, logger=None):
    """dist.py:30-67"""
    backend = getattr(args, "dist_backend", None) or ("nccl" if torch.cuda.is_available() else "gloo")
    timeout = getattr(args, "timeout", 86400)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group(backend=backend, init_method="env://", timeout=datetime.timedelta(seconds=timeout))
        env = DistEnv(dist.get_world_size(), dist.get_rank(), local, 1, dist.get_rank() == 0,
                      torch.cuda.get_device_name() if torch.cuda.is_available() else "cpu")
    else:
        env = DistEnv(1, 0, 0, torch.cuda.device_count(), True,
                      torch.cuda.get_device_name() if torch.cuda.is_available() else "cpu")
    if logger is not None:
        logger.info(env)
    return env


class _ModuleBox(torch.nn.Module):
    """what the callers need from DDP / DataParallel: ``.module`` (main_sampling_fid.py:210 bypasses DDP.forward)"""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


def dataparallel_and_sync(distenv, model, find_unused_parameters=True):
    """dist.py:70-85 -- replicate rank 0's weights, return a wrapper with .module"""
    if dist.is_initialized() and distenv.world_size > 1:
        tensors = [t for t in model.state_dict().values()]
        by_dtype = {}
        for t in tensors:
            by_dtype.setdefault((t.dtype, t.device), []).append(t)
        for (dt, dev), ts in by_dtype.items():
            flat = torch.cat([t.reshape(-1) for t in ts])
            dist.broadcast(flat, 0)
            off = 0
            for t in ts:
                t.copy_(flat[off:off + t.numel()].view_as(t))
                off += t.numel()
        if hasattr(model, "_invalidate_native"):
            model._invalidate_native()
        dist.barrier()
    return _ModuleBox(model)


def all_gather_cat(distenv, tensor, dim=0):
    """dist.py:94-103"""
    if distenv.world_size == 1 or not dist.is_initialized():
        return tensor
    out = [torch.empty_like(tensor) for _ in range(distenv.world_size)]
    dist.all_gather(out, tensor.contiguous())
    return torch.cat(out, dim=dim)


def shard_batch(distenv, total, rank=None):
    """[lo, hi) slice of a global batch of independent images owned by this rank (main_sampling_fid.py:196-206)"""
    r = distenv.world_rank if rank is None else rank
    per = (total + distenv.world_size - 1) // distenv.world_size
    lo = min(r * per, total)
    return lo, min(lo + per, total)
