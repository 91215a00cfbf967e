"""NormalGamma node — independent Normal-Gamma per feature, "no matrix inversions required" — with the reference's
interface (dists/NormalGamma.py:4-121).  The O(N K d) passes run in libvbmp_b200.so: Elog_like and the
responsibilities of a GaussianMixtureModel(isotropic=True) in the streaming diagonal E-step (vbmp_diag_estep), the weighted
statistics sum r x, sum r x^2, sum r as the diagonal mode of the Gram kernel (vbmp_gram flags bit 1: 2 d + 1 of the pair
columns, HBM-bound); the K x d update is element-wise torch on the device, in the reference's operation order."""
from __future__ import annotations

import math

import torch

from . import _lib, _shapes
from .gamma import Gamma


class NormalGamma():
    def __init__(self, event_shape, batch_shape=(), scale=torch.tensor(1.0),
                 prior_parms={'lambda_mu': torch.tensor(1.0), 'mu': torch.tensor(0.0),
                              'alpha': torch.tensor(2.0), 'beta': torch.tensor(2.0)}):
        """dists/NormalGamma.py:6-28 (same RNG draws, in order: rand_like(lambda), the Gamma node's two rand, randn_like(mu))."""
        self.dim = event_shape[-1]
        self.event_dim = 1                                      # (sic: dists/NormalGamma.py:13)
        self.event_shape = event_shape
        self.batch_dim = len(batch_shape)
        self.batch_shape = batch_shape
        dev = torch.empty(0).device
        self.lambda_mu_0 = prior_parms['lambda_mu'].to(dev).expand(batch_shape + event_shape[:-1])
        self.lambda_mu = self.lambda_mu_0 + torch.rand_like(self.lambda_mu_0, requires_grad=False)
        self.mu_0 = prior_parms['mu'].to(dev).expand(batch_shape + event_shape)
        self.gamma = Gamma(event_shape=event_shape, batch_shape=batch_shape,
                           prior_parms={'alpha': prior_parms['alpha'].to(dev), 'beta': prior_parms['beta'].to(dev) * scale ** 2})
        self.mu = self.mu_0 + torch.randn_like(self.mu_0, requires_grad=False) / self.gamma.mean().sqrt()
        self.SExx = 0.0
        self.SEx = 0.0
        self.N = 0.0

    def to_event(self, n):
        if n == 0:
            return self
        self.event_dim = self.event_dim + n
        self.batch_dim = self.batch_dim - n
        self.event_shape = self.batch_shape[-n:] + self.event_shape
        self.batch_shape = self.batch_shape[:-n]
        self.gamma.to_event(n)
        return self

    def to(self, device):
        for k in ("lambda_mu_0", "lambda_mu", "mu_0", "mu", "SExx", "SEx", "N"):
            if isinstance(getattr(self, k), torch.Tensor):
                setattr(self, k, getattr(self, k).to(device))
        self.gamma.to(device)
        return self

    # ---- helpers ------------------------------------------------------------------------------------
    def _full(self):
        full = tuple(self.batch_shape) + tuple(self.event_shape[:-1])
        return full, int(math.prod(full))

    def _plan(self, X):
        nb, ne = self.batch_dim, self.event_dim
        sample_shape = tuple(X.shape[:X.ndim - nb - ne])
        bstar = tuple(X.shape[X.ndim - nb - ne:X.ndim - ne])
        return _shapes.make_plan(self.batch_shape, self.event_shape[:-1], bstar, sample_shape)

    def _prep(self, plan, logprior=None):
        """(mu, tau, cst) in kernel order (G*K, d): tau = gamma.mean(), cst = 1/2 sum_i gamma.loggeomean()_i [+ log prior]."""
        nb, nx, d, dev = self.batch_dim, self.event_dim - 1, self.dim, self.mu.device
        full, C = self._full()
        f = _lib.f32
        tk = lambda t, tail: _shapes.theta_to_GK(t, plan, nb, nx, tail)   # noqa: E731
        mu = tk(f(self.mu.expand(full + (d,)), dev), 1)
        tau = tk(f(self.gamma.mean().expand(full + (d,)), dev), 1)
        cst = 0.5 * self.gamma.loggeomean().sum(-1)
        if logprior is not None:
            cst = cst + logprior.view(tuple(logprior.shape) + (cst.ndim - logprior.ndim) * (1,))
        cst = tk(f(cst.expand(full), dev), 0)
        return mu, tau, cst

    # ---- reference protocol ---------------------------------------------------------------------------
    def ss_update(self, SExx, SEx, N, lr=1.0, beta=None):
        """dists/NormalGamma.py:41-56 (K x d element-wise, the reference's operation order)."""
        if beta is not None:
            self.SExx = SExx + beta * self.SExx
            self.SEx = SEx + beta * self.SEx
            self.N = N + beta * self.N
            SExx = self.SExx
            SEx = self.SEx
            N = self.N
        lambda_mu = self.lambda_mu_0 + N
        mu = (self.lambda_mu_0.unsqueeze(-1) * self.mu_0 + SEx) / lambda_mu.unsqueeze(-1)
        SExx = SExx + self.lambda_mu_0.unsqueeze(-1) * self.mu_0 ** 2 - lambda_mu.unsqueeze(-1) * mu ** 2
        self.lambda_mu = lr * lambda_mu + (1 - lr) * self.lambda_mu
        self.mu = lr * mu + (1 - lr) * self.mu
        self.gamma.ss_update(0.5 * N.unsqueeze(-1), 0.5 * SExx, lr, beta)

    def _stats(self, X, p=None):
        """sum r x^2, sum r x, sum r from ONE weighted pass (the Gram kernel's diagonal mode)."""
        plan = self._plan(X)
        dev, d = self.mu.device, self.dim
        Xc = _lib.f32(X, dev).reshape(plan.N, plan.GX, d)
        pc = None if p is None else _lib.f32(p, dev).reshape(plan.N, plan.GP, plan.K)
        xg = _shapes.idx_tensor(plan.xg, dev)
        pg = _shapes.idx_tensor(plan.pg, dev)
        G = _lib.gram(Xc, None, plan.N, plan.GX, xg, pc, plan.GP, pg, plan.G, plan.K, _lib.pad_dim(d), diag=True)
        G = _shapes.GK_to_theta(G, plan, (d + 1, d + 1))
        SExx = G[..., :d, :d].diagonal(dim1=-2, dim2=-1)
        SEx = G[..., :d, d]
        if p is None:
            N = torch.tensor(float(plan.N), device=dev).expand(self.batch_shape + self.event_shape[:-1])
        else:
            N = G[..., d, d]
            for _ in range(self.event_dim - 1):                  # the reference's N is p summed over samples: batch dims only
                N = N[..., 0]
        return SExx, SEx, N

    def raw_update(self, X, p=None, lr=1.0, beta=None):
        """dists/NormalGamma.py:58-73."""
        SExx, SEx, N = self._stats(X, p)
        self.ss_update(SExx, SEx, N, lr, beta)

    def Elog_like(self, X):
        """dists/NormalGamma.py:76-86 (the expression of :83) -> vbmp_diag_estep, mode 0."""
        plan = self._plan(X)
        dev = self.mu.device
        mu, tau, cst = self._prep(plan)
        Xc = _lib.f32(X, dev).reshape(plan.N, plan.GX, self.dim)
        out = _lib.diag_estep(Xc, plan.N, plan.GX, _shapes.idx_tensor(plan.xg, dev), mu, tau, cst, plan.G, plan.K, self.dim, 0)
        out = _shapes.logits_to_ref(out, plan)
        for i in range(self.event_dim - 1):
            out = out.sum(-1)
        return out

    def KLqprior(self):
        """dists/NormalGamma.py:88-94."""
        out = self.lambda_mu_0 / 2.0 * ((self.mu - self.mu_0) ** 2 * self.gamma.mean()).sum(-1)
        out = out + self.dim / 2.0 * (self.lambda_mu_0 / self.lambda_mu - (self.lambda_mu_0 / self.lambda_mu).log() - 1)
        for i in range(self.event_dim - 1):
            out = out.sum(-1)
        return out + self.gamma.KLqprior().sum(-1)

    def mean(self):
        return self.mu

    def Emumu(self):
        return self.mu.unsqueeze(-2) * self.mu.unsqueeze(-1) + self.ESigma() / self.lambda_mu.unsqueeze(-1).unsqueeze(-1)

    def ElogdetinvSigma(self):
        return self.gamma.loggeomean().sum(-1)

    def EmuTinvSigmamu(self):
        return (self.mu ** 2 * self.gamma.mean()).sum(-1) + self.dim / self.lambda_mu

    def EXTinvUX(self):
        return (self.mu ** 2 * self.gamma.mean()).sum(-1) + self.dim / self.lambda_mu

    def EinvSigma(self):
        return self.gamma.mean().unsqueeze(-1) * torch.eye(self.dim, requires_grad=False, device=self.mu.device)

    def ESigma(self):
        return self.gamma.meaninv().unsqueeze(-1) * torch.eye(self.dim, requires_grad=False, device=self.mu.device)

    def Res(self):
        return -0.5 * self.EXTinvUX() + 0.5 * self.ElogdetinvSigma() - 0.5 * self.dim * math.log(2 * math.pi)

    def EinvSigmamu(self):
        return self.gamma.mean() * self.mu
