"""Wishart node with the reference's interface (dists/Wishart.py:7-97), state as plain tensor
attributes, arithmetic in libvbmp_b200.so (batched Cholesky / inverse / logdet / multivariate
digamma+lgamma kernels)."""
from __future__ import annotations

import math

import torch

from . import _lib


def _bcast_N(N, shape, device):
    N = torch.as_tensor(N, dtype=torch.float32, device=device)
    return _lib.f32(N.expand(shape)).reshape(-1)


class Wishart():
    def __init__(self, event_shape, batch_shape=(), scale=torch.tensor(1.0, requires_grad=False)):
        """dists/Wishart.py:9-26: invU_0 = scale^2 I (stride-0 expanded), nu_0 = dim + 2."""
        assert event_shape[-1] == event_shape[-2]
        self.dim = event_shape[-1]
        self.event_shape = event_shape
        self.event_dim = len(event_shape)
        self.batch_dim = len(batch_shape)
        self.batch_shape = batch_shape

        self.invU_0 = (scale ** 2 * torch.eye(self.dim, requires_grad=False)).expand(batch_shape + event_shape)
        self.nu_0 = torch.tensor(self.dim + 2.0).expand(batch_shape + event_shape[:-2])
        # constructor-time constants of a scaled identity (one-off, closed form: logdet = dim log scale^2, U = I / scale^2);
        # scale may be any tensor that broadcasts against batch_shape + event_shape, as in the reference
        s2 = torch.as_tensor(scale, dtype=torch.float32) ** 2
        if s2.ndim >= 2:
            assert s2.shape[-1] == 1 and s2.shape[-2] == 1, "scale must broadcast as a scalar per matrix"
            ld = self.dim * s2[..., 0, 0].log()
        else:
            assert s2.numel() == 1, "scale must broadcast as a scalar per matrix"
            ld = self.dim * s2.reshape(()).log()
        self.logdet_invU_0 = ld.expand(tuple(batch_shape + event_shape[:-2])).clone()
        self.invU = self.invU_0
        self.U = (torch.eye(self.dim, requires_grad=False) / s2).expand(batch_shape + event_shape)
        self.nu = self.nu_0
        self.logdet_invU = self.logdet_invU_0.clone()
        self.SExx = 0.0
        self.N = 0.0
        self.info = None      # device int32 per-matrix Cholesky status of the last update (0 = SPD)

    def to_event(self, n):
        """dists/Wishart.py:28-35."""
        if n == 0:
            return self
        self.event_dim = self.event_dim + n
        self.batch_dim = self.batch_dim - n
        self.event_shape = self.batch_shape[-n:] + self.event_shape
        self.batch_shape = self.batch_shape[:-n]
        return self

    def to(self, device):
        for k in ("invU_0", "nu_0", "logdet_invU_0", "invU", "U", "nu", "logdet_invU"):
            setattr(self, k, getattr(self, k).to(device))
        for k in ("SExx", "N"):
            if isinstance(getattr(self, k), torch.Tensor):
                setattr(self, k, getattr(self, k).to(device))
        return self

    # ---- kernels ----------------------------------------------------------------------------------
    def _mat_shape(self):
        return tuple(self.invU_0.shape)

    def _flat(self):
        """(C, d, d) / (C,) contiguous fp32 views of the state."""
        ms = self._mat_shape()
        dev = self.invU.device
        C = int(math.prod(ms[:-2]))
        d = self.dim
        g = lambda t, s: _lib.f32(t.expand(s), dev).reshape((C,) + tuple(s[len(ms) - 2:]))   # noqa: E731
        return C, d, g, ms

    def log_mvgamma(self, nu):
        """dists/Wishart.py:37-38 (K x d values; plain torch on the node's device)."""
        return (nu.unsqueeze(-1) - torch.arange(self.dim, device=nu.device) / 2.0).lgamma().sum(-1)

    def log_mvdigamma(self, nu):
        """dists/Wishart.py:40-41."""
        return (nu.unsqueeze(-1) - torch.arange(self.dim, device=nu.device) / 2.0).digamma().sum(-1)

    def ss_update(self, SExx, N, lr=1.0, beta=None):
        """dists/Wishart.py:43-56 -> vbmp_wishart_update."""
        assert (SExx.ndim == self.batch_dim + self.event_dim)
        assert (N.ndim == self.batch_dim + self.event_dim - 2)
        if beta is not None:
            self.SExx = SExx + beta * self.SExx
            self.N = N + beta * self.N
            SExx = self.SExx
            N = self.N
        C, d, g, ms = self._flat()
        dev = self.invU.device
        invU, nu, U, logdet, info = _lib.wishart_update(
            g(SExx, ms), _bcast_N(N, ms[:-2], dev), g(self.invU_0, ms), g(self.nu_0, ms[:-2]),
            g(self.invU, ms), g(self.nu, ms[:-2]), C, d, float(lr))
        self._set(invU, nu, U, logdet, info)

    def _set(self, invU, nu, U, logdet, info):
        ms = self._mat_shape()
        self.invU = invU.view(ms)
        self.U = U.view(ms)
        self.nu = nu.view(ms[:-2])
        self.logdet_invU = logdet.view(ms[:-2])
        self.info = info

    def check(self):
        """Raise if the last update hit a non-SPD matrix (synchronises)."""
        if self.info is not None and bool((self.info != 0).any()):
            bad = int((self.info != 0).nonzero()[0])
            raise _lib.VbmpError(f"Wishart update: matrix {bad} is not positive definite (info={int(self.info[bad])})")

    def mean(self):
        return self.U * self.nu.view(self.nu.shape + (1, 1))

    def meaninv(self):
        return self.invU / (self.nu.view(self.nu.shape + (1, 1)) - self.dim - 1)

    def ESigma(self):
        return self.invU / (self.nu.view(self.nu.shape + (1, 1)) - self.dim - 1)

    def EinvSigma(self):
        return self.U * self.nu.view(self.nu.shape + (1, 1))

    def invEinvSigma(self):
        return self.invU / (self.nu.view(self.nu.shape + (1, 1)))

    def ElogdetinvSigma(self):
        """dists/Wishart.py:82-83 -> vbmp_wishart_elogdet (CUDA only, like every other kernel-backed method: a node on the
        CPU raises VbmpError; the class install() binds keeps the reference's own code for CPU-resident nodes)."""
        C, d, g, ms = self._flat()
        out = _lib.wishart_elogdet(g(self.nu, ms[:-2]), g(self.logdet_invU, ms[:-2]), C, d)
        return out.view(ms[:-2])

    def logdetEinvSigma(self):
        return -self.logdet_invU + self.nu.log()

    def KLqprior(self):
        """dists/Wishart.py:88-94 -> vbmp_wishart_kl."""
        C, d, g, ms = self._flat()
        out = _lib.wishart_kl(g(self.invU_0, ms), g(self.U, ms), g(self.nu_0, ms[:-2]), g(self.nu, ms[:-2]),
                              g(self.logdet_invU, ms[:-2]), g(self.logdet_invU_0, ms[:-2]), C, d).view(ms[:-2])
        for i in range(self.event_dim - 2):
            out = out.sum(-1)
        return out

    def logZ(self):
        return self.log_mvgamma(self.nu / 2.0) + 0.5 * self.nu * self.dim * math.log(2.0) \
            - 0.5 * self.nu * self.logdet_invU
