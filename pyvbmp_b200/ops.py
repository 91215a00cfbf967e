"""torch custom ops over the C ABI (`torch.ops.vbmp.*`): the "thin torch custom-op layer" of the north star.

The mirrors in this package call the ctypes binding (`_lib.py`) directly; these registrations expose the same four hot
entry points to code that wants dispatcher-visible ops (shape inference under FakeTensor / `torch.compile` graphs around
the EM loop, profiler names).  CUDA only: there is no CPU kernel to register, a CPU tensor raises from the binding.

    torch.ops.vbmp.estep_logits(z0, z1, W, m, cst)              -> logits (N, K)
    torch.ops.vbmp.estep_assign(z0, z1, W, m, cst)              -> (p (N, K), logZn (N,), NA (K,), logZ ())
    torch.ops.vbmp.gram(z0, z1, p, diag)                        -> (K, D+1, D+1)
    torch.ops.vbmp.hmm_forward_backward(logits, trans, init)    -> (p (T,S,K), SEzz (S,K,K), SEz0 (S,K), logZ (S,))

z0 (N, d0), z1 (N, d1) or None, W (K, Dp, Dp), m (K, Dp), cst (K,) as produced by vbmp_niw_prep / vbmp_mnw_prep (`_lib.niw_prep`).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib, _shapes


def _xg(dev):
    return _shapes.idx_tensor((0,), dev)


@torch.library.custom_op("vbmp::estep_logits", mutates_args=())
def estep_logits(z0: torch.Tensor, z1: Optional[torch.Tensor], W: torch.Tensor, m: torch.Tensor, cst: torch.Tensor) -> torch.Tensor:
    N, K, Dp = z0.shape[0], cst.shape[0], W.shape[-1]
    z1v = None if z1 is None else _lib.f32(z1).view(N, 1, -1)
    out = _lib.estep(_lib.f32(z0).view(N, 1, -1), z1v, N, 1, _xg(z0.device), W, m, cst, 1, K, Dp, 0)
    return out.view(N, K)


@estep_logits.register_fake
def _(z0, z1, W, m, cst):
    return z0.new_empty((z0.shape[0], cst.shape[0]))


@torch.library.custom_op("vbmp::estep_assign", mutates_args=())
def estep_assign(z0: torch.Tensor, z1: Optional[torch.Tensor], W: torch.Tensor, m: torch.Tensor,
                 cst: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    N, K, Dp = z0.shape[0], cst.shape[0], W.shape[-1]
    z1v = None if z1 is None else _lib.f32(z1).view(N, 1, -1)
    p, lzn, NA, lZ = _lib.estep(_lib.f32(z0).view(N, 1, -1), z1v, N, 1, _xg(z0.device), W, m, cst, 1, K, Dp, 1)
    return p.view(N, K), lzn.view(N), NA.view(K), lZ.view(())


@estep_assign.register_fake
def _(z0, z1, W, m, cst):
    N, K = z0.shape[0], cst.shape[0]
    return z0.new_empty((N, K)), z0.new_empty((N,)), z0.new_empty((K,)), z0.new_empty(())


@torch.library.custom_op("vbmp::gram", mutates_args=())
def gram(z0: torch.Tensor, z1: Optional[torch.Tensor], p: Optional[torch.Tensor], diag: bool = False) -> torch.Tensor:
    N = z0.shape[0]
    K = 1 if p is None else p.shape[-1]
    d1 = 0 if z1 is None else z1.shape[-1]
    D = z0.shape[-1] + d1
    z1v = None if z1 is None else _lib.f32(z1).view(N, 1, -1)
    pv = None if p is None else _lib.f32(p).view(N, 1, K)
    xg = _xg(z0.device)
    G = _lib.gram(_lib.f32(z0).view(N, 1, -1), z1v, N, 1, xg, pv, 1, xg, 1, K, _lib.pad_dim(D), diag=diag)
    return G.view(K, D + 1, D + 1)


@gram.register_fake
def _(z0, z1, p, diag=False):
    K = 1 if p is None else p.shape[-1]
    D = z0.shape[-1] + (0 if z1 is None else z1.shape[-1])
    return z0.new_empty((K, D + 1, D + 1))


@torch.library.custom_op("vbmp::hmm_forward_backward", mutates_args=())
def hmm_forward_backward(logits: torch.Tensor, trans: torch.Tensor, init: torch.Tensor,
                         ptemp: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    T, S, K = logits.shape
    return _lib.hmm_forward_backward(_lib.f32(logits), _lib.f32(trans).view(1, K, K), _lib.f32(init).view(1, K), T, S, 1, K, ptemp)


@hmm_forward_backward.register_fake
def _(logits, trans, init, ptemp=1.0):
    T, S, K = logits.shape
    return logits.new_empty((T, S, K)), logits.new_empty((S, K, K)), logits.new_empty((S, K)), logits.new_empty((S,))
