"""torch custom ops over the C ABI (`torch.ops.vbmp.*`): the "thin torch custom-op layer" of the north star.

The mirrors in this package call the ctypes binding (`_lib.py`) directly; these registrations expose the same four hot
entry points to code that wants dispatcher-visible ops (shape inference under FakeTensor / `torch.compile` graphs around
the EM loop, profiler names); the widened rows are registered too.  CUDA only: there is no CPU kernel to register, a CPU tensor raises from the binding.

    torch.ops.vbmp.estep_logits(z0, z1, W, m, cst)              -> logits (N, K)
    torch.ops.vbmp.estep_assign(z0, z1, W, m, cst)              -> (p (N, K), logZn (N,), NA (K,), logZ ())
    torch.ops.vbmp.gram(z0, z1, p, diag)                        -> (K, D+1, D+1)
    torch.ops.vbmp.hmm_forward_backward(logits, trans, init)    -> (p (T,S,K), SEzz (S,K,K), SEz0 (S,K), logZ (S,))
    torch.ops.vbmp.rowgemm(A, B, bias)                          -> A (N, Kd) @ B (Kd, M) (+ bias)          (predict's shared-operand products)
    torch.ops.vbmp.rowterm(A, B, C, alpha)                      -> C (N, K) + alpha A (N, F) @ B (F, K)    (covariance trace terms)
    torch.ops.vbmp.wsum(p, S)                                   -> p (N, K)^T @ S (N, F)                   (weighted covariance sums)
    torch.ops.vbmp.moe_moments(mean, p, base)                   -> (mu (N, n), Sigma (N, n, n))            (mixture-of-experts moments)

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


@torch.library.custom_op("vbmp::rowgemm", mutates_args=())
def rowgemm(A: torch.Tensor, B: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _lib.rowgemm(_lib.f32(A), _lib.f32(B), bias=None if bias is None else _lib.f32(bias))


@rowgemm.register_fake
def _(A, B, bias=None):
    return A.new_empty((A.shape[0], B.shape[1]))


@torch.library.custom_op("vbmp::rowterm", mutates_args=())
def rowterm(A: torch.Tensor, B: torch.Tensor, C: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    A, B = _lib.f32(A), _lib.f32(B)
    out = _lib.f32(C).clone()
    N, F = A.shape
    K = B.shape[1]
    if _lib.rowterm_supported(N, F, K, A.stride(0)):
        return _lib.rowterm(A, B, C=out, alpha=alpha, accumulate=True)
    return out.add_(_lib.rowgemm(A, B), alpha=alpha)


@rowterm.register_fake
def _(A, B, C, alpha=1.0):
    return C.new_empty(C.shape)


@torch.library.custom_op("vbmp::wsum", mutates_args=())
def wsum(p: torch.Tensor, S: torch.Tensor) -> torch.Tensor:
    p, S = _lib.f32(p), _lib.f32(S)
    if _lib.wsum_supported(p.shape[0], p.shape[1], S.shape[1], S.stride(0)):
        return _lib.wsum(p, S)
    return _lib.rowgemm(p.t().contiguous(), S)                  # shapes outside the kernel's window


@wsum.register_fake
def _(p, S):
    return p.new_empty((p.shape[1], S.shape[1]))


@torch.library.custom_op("vbmp::moe_moments", mutates_args=())
def moe_moments(mean: torch.Tensor, p: torch.Tensor, base: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    N, K, n = mean.shape
    return _lib.moe_moments(_lib.f32(mean), _lib.f32(p), None if base is None else _lib.f32(base), N, K, n)


@moe_moments.register_fake
def _(mean, p, base=None):
    N, K, n = mean.shape
    return mean.new_empty((N, n)), mean.new_empty((N, n, n))
