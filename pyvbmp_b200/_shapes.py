"""Flatten the reference's ``sample_shape + batch_shape* + event_shape`` convention (SURVEY.md §8b,
Appendix E) into the kernel layout: theta groups G x mixture axis K, data columns GX, weights GP.

Pure Python / shape arithmetic (runs on CPU, unit-tested without a GPU).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Tuple

import torch


@dataclass
class Plan:
    sample_shape: Tuple[int, ...]
    N: int
    lead: Tuple[int, ...]        # batch dims treated as theta groups (all but the mixture axis)
    K: int                       # mixture axis (last batch dim when the data broadcasts over it, else 1)
    extra: Tuple[int, ...]       # extra event dims (event_shape minus the trailing feature dims)
    G: int                       # theta groups = prod(lead) * prod(extra)
    GX: int                      # distinct data columns
    xg: Tuple[int, ...]          # theta group -> data column
    GP: int                      # distinct weight columns (= prod(lead))
    pg: Tuple[int, ...]          # theta group -> weight column
    k_is_batch: bool             # True when K is the last batch dim


def prod(s):
    return int(math.prod(s))


def make_plan(batch_shape, extra_event, data_batch_star, sample_shape) -> Plan:
    """``data_batch_star``: the data's sizes along the batch dims (1 where it broadcasts)."""
    batch_shape, extra_event = tuple(batch_shape), tuple(extra_event)
    bstar = tuple(data_batch_star)
    assert len(bstar) == len(batch_shape), (bstar, batch_shape)
    for a, b in zip(bstar, batch_shape):
        assert a in (1, b), f"data batch dims {bstar} do not broadcast against batch_shape {batch_shape}"
    if len(batch_shape) >= 1 and bstar[-1] == 1:
        lead, K, lead_star, k_is_batch = batch_shape[:-1], batch_shape[-1], bstar[:-1], True
    else:   # every component sees its own data column: no shared mixture axis
        lead, K, lead_star, k_is_batch = batch_shape, 1, bstar, False
    G = prod(lead) * prod(extra_event)
    GX = prod(lead_star) * prod(extra_event)
    GP = prod(lead)
    xg, pg = [], []
    nl = len(lead)
    for g in range(G):
        e = g % max(prod(extra_event), 1)
        b = g // max(prod(extra_event), 1)
        # unravel b over lead, re-ravel over lead_star with broadcasting
        idx, rem = [], b
        for s in reversed(lead):
            idx.append(rem % s)
            rem //= s
        idx = idx[::-1]
        bx = 0
        for i in range(nl):
            bx = bx * lead_star[i] + (idx[i] if lead_star[i] != 1 else 0)
        xg.append(bx * max(prod(extra_event), 1) + e)
        pg.append(b)
    return Plan(tuple(sample_shape), prod(sample_shape), lead, K, extra_event, G, GX, tuple(xg), GP, tuple(pg),
                k_is_batch)


def theta_to_GK(t, plan: Plan, batch_dim, n_extra, tail):
    """Parameter tensor of shape batch + extra + tail -> (G*K, *tail) contiguous, group-major."""
    bs = t.shape[:batch_dim]
    ex = t.shape[batch_dim:batch_dim + n_extra]
    tl = tuple(t.shape[batch_dim + n_extra:])
    assert len(tl) == tail, (t.shape, batch_dim, n_extra, tail)
    if plan.k_is_batch:
        # (lead..., K, extra..., tail) -> (lead..., extra..., K, tail)
        nd = t.ndim
        perm = list(range(batch_dim - 1)) + list(range(batch_dim, batch_dim + n_extra)) + [batch_dim - 1] \
            + list(range(batch_dim + n_extra, nd))
        t = t.permute(perm)
    return t.reshape((plan.G * plan.K,) + tl).contiguous()


def GK_to_theta(t, plan: Plan, tail_shape):
    """(G*K, *tail) or (G, K, *tail) -> batch + extra + tail (inverse of theta_to_GK)."""
    tail_shape = tuple(tail_shape)
    t = t.reshape(plan.lead + plan.extra + (plan.K,) + tail_shape)
    nl, ne = len(plan.lead), len(plan.extra)
    if plan.k_is_batch:
        perm = list(range(nl)) + [nl + ne] + list(range(nl, nl + ne)) + list(range(nl + ne + 1, t.ndim))
        return t.permute(perm)
    return t.reshape(plan.lead + plan.extra + tail_shape)


def data_to_cols(X, plan: Plan, batch_dim, n_extra, feat_dims=1):
    """Data of shape sample + batch* + extra + feature dims -> (N, GX, d) contiguous fp32."""
    d = prod(X.shape[X.ndim - feat_dims:])
    return X.reshape(plan.N, plan.GX, d)


def logits_to_ref(out, plan: Plan):
    """Kernel output (N, G, K) -> sample + batch (+ extra dims still present, to be summed by the caller)."""
    t = out.reshape(plan.sample_shape + plan.lead + plan.extra + (plan.K,))
    ns, nl, ne = len(plan.sample_shape), len(plan.lead), len(plan.extra)
    if plan.k_is_batch:
        perm = list(range(ns + nl)) + [ns + nl + ne] + list(range(ns + nl, ns + nl + ne))
        t = t.permute(perm)
    else:
        t = t.reshape(plan.sample_shape + plan.lead + plan.extra)
    return t


_IDX_CACHE = {}


def idx_tensor(v, device):
    """int32 index vector on `device` (the kernels' xg / pg group maps).  Cached: the same few tuples are asked for by every
    E-step / Gram call, and building one is a pageable host -> device copy (~15 us of host time each, 2-3 per EM iteration,
    two per streamed chunk).  The kernels only read them."""
    key = (tuple(int(i) for i in v), str(device))
    t = _IDX_CACHE.get(key)
    if t is None:
        if len(_IDX_CACHE) > 256:
            _IDX_CACHE.clear()
        t = _IDX_CACHE[key] = torch.tensor(list(key[0]), dtype=torch.int32, device=device)
    return t
