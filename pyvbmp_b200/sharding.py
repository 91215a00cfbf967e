"""Sample sharding across GPUs: each rank owns a slice of the rows, parameters are
replicated, and ONE all-reduce per EM iteration sums the packed block
[Gram statistics | logZ | NA] before the identical replicated update on every rank.
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) is the plumbing.
"""
from __future__ import annotations

import torch

_state = {"group": None, "enabled": False}


def enable(group=None):
    """Turn on sample sharding for Mixture / MixtureofLinearTransforms updates in this process."""
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _state["group"] = group
    _state["enabled"] = dist.get_world_size(group) > 1


def disable():
    _state["group"] = None
    _state["enabled"] = False


def enabled():
    return _state["enabled"]


def shard_rows(n_total, rank, world):
    """Contiguous row range [lo, hi) owned by ``rank`` (remainder spread over the first ranks)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_packed(tensors):
    """Sum a list of tensors across ranks with a single collective; returns new tensors."""
    import torch.distributed as dist
    flat = torch.cat([t.reshape(-1).to(torch.float32) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=_state["group"])
    out, o = [], 0
    for t in tensors:
        n = t.numel()
        out.append(flat[o:o + n].view(t.shape))
        o += n
    return out


def broadcast_(t, src=0):
    import torch.distributed as dist
    dist.broadcast(t, src=src, group=_state["group"])
    return t
