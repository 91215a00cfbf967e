"""Mixture / GaussianMixtureModel with the reference's interface (dists/Mixture.py:5-127,
models/GaussianMixtureModel.py:6-16).  The EM loop stays host Python; update_assignments is ONE fused
kernel sequence (K1 prep + K2 E-step with the logsumexp / responsibilities / NA / logZ epilogue) when
the mixture axis is the last batch dim of a NormalInverseWishart with a vector event.
"""
from __future__ import annotations

import torch

from . import _lib, _shapes, sharding
from .dirichlet import Dirichlet
from .niw import NormalInverseWishart


def fused_update_assignments(self, X):
    """dists/Mixture.py:38-45 on the CUDA path.  Also used as the method patch install() puts on the
    reference's Mixture class."""
    dist = self.dist
    fusable = (isinstance(dist, NormalInverseWishart) and self.event_dim == 1 and dist.event_dim == 1)
    if not fusable:
        return generic_update_assignments(self, X)
    Xv = X.view(X.shape[:-dist.event_dim] + self.event_dim * (1,) + dist.event_shape)
    plan = dist._plan(Xv)
    if not plan.k_is_batch:
        return generic_update_assignments(self, X)
    dev = dist.mu.device
    W, m, cst, info, Dp = dist._prep(plan, logprior=self.pi.loggeomean())
    Xc = _lib.f32(Xv, dev).reshape(plan.N, plan.GX, dist.dim)
    p, logZn, NA, logZ = _lib.estep(Xc, None, plan.N, plan.GX, _shapes.idx_tensor(plan.xg, dev), W, m, cst,
                                    plan.G, plan.K, Dp, 1)
    self.p = p.view(plan.sample_shape + plan.lead + (plan.K,))
    self.logZ_n = logZn.view(plan.sample_shape + plan.lead)
    self.NA = NA.view(plan.lead + (plan.K,))
    self.logZ = logZ.view(plan.lead)


def generic_update_assignments(self, X):
    """Any other event / batch structure: logits from dist.Elog_like (CUDA), softmax glue in torch."""
    log_p = self.Elog_like(X)
    dims = list(range(-self.event_dim, 0))
    logZ = torch.logsumexp(log_p, dim=dims)
    self.p = (log_p - logZ.view(logZ.shape + self.event_dim * (1,))).exp()
    sample_dim = self.p.ndim - self.batch_dim - self.event_dim
    self.NA = self.p.sum(list(range(sample_dim)))
    self.logZ = logZ.sum(list(range(sample_dim)))


class Mixture():

    def __init__(self, dist, event_shape, prior_parms={'alpha': torch.tensor(0.5)}):
        """dists/Mixture.py:8-19."""
        assert dist.batch_shape[-len(event_shape):] == event_shape
        self.event_shape = event_shape
        self.event_dim = len(event_shape)
        self.batch_shape = dist.batch_shape[:-len(event_shape)]
        self.batch_dim = len(self.batch_shape)

        self.pi = Dirichlet(event_shape=event_shape, batch_shape=self.batch_shape, prior_parms=prior_parms)
        self.dist = dist
        self.logZ = torch.tensor(-torch.inf, requires_grad=False)
        self.ELBO_last = torch.tensor(-torch.inf)

    def to_event(self, n):
        if n == 0:
            return self
        self.event_dim = self.event_dim + n
        self.event_shape = self.batch_shape[-n:] + self.event_shape
        self.batch_shape = self.batch_shape[:-n]
        self.pi.to_event(n)
        self.dist.to_event(n)
        return self

    def to(self, device):
        self.pi.to(device)
        self.dist.to(device)
        self.logZ = self.logZ.to(device)
        self.ELBO_last = self.ELBO_last.to(device)
        return self

    update_assignments = fused_update_assignments

    def update_parms(self, X, lr=1.0):
        """dists/Mixture.py:47-49."""
        self.pi.ss_update(self.NA, lr=lr)
        self.update_dist(X, lr=lr)

    def raw_update(self, X, iters=1, lr=1.0, verbose=False):
        self.update(X, iters=iters, lr=lr, verbose=verbose)

    def update(self, X, iters=1, lr=1.0, verbose=False):
        """dists/Mixture.py:54-62: E-step, ELBO with pre-M-step parameters, M-step."""
        for i in range(iters):
            self.update_assignments(X)
            if sharding.enabled():
                ELBO = self._sharded_m_step(X, lr)
            else:
                ELBO = self.ELBO()
                self.update_parms(X, lr)
            if verbose:
                print('Percent Change in ELBO:   ', (ELBO - self.ELBO_last) / self.ELBO_last.abs() * 100.0)
            self.ELBO_last = ELBO

    def _sharded_m_step(self, X, lr):
        """Rows are sharded over ranks: local Gram + ONE all-reduce of [Gram | logZ | NA], then the
        replicated update (identical on every rank).  ELBO still uses the pre-update parameters."""
        Xv = X.view(X.shape[:-self.dist.event_dim] + self.event_dim * (1,) + self.dist.event_shape)
        G, plan = self.dist._gram(Xv, self.p)
        G, logZ, NA = sharding.all_reduce_packed([G, self.logZ, self.NA])
        self.logZ, self.NA = logZ, NA
        ELBO = self.ELBO()
        self.pi.ss_update(self.NA, lr=lr)
        self.dist._update_from_gram(G, plan, True, lr)
        return ELBO

    def update_dist(self, X, lr):
        """dists/Mixture.py:64-66."""
        Xv = X.view(X.shape[:-self.dist.event_dim] + self.event_dim * (1,) + self.dist.event_shape)
        self.dist.raw_update(Xv, self.p, lr)

    def Elog_like(self, X):
        """dists/Mixture.py:68-70."""
        X = X.view(X.shape[:-self.dist.event_dim] + self.event_dim * (1,) + self.dist.event_shape)
        return self.dist.Elog_like(X) + self.pi.loggeomean()

    def KLqprior(self):
        return self.dist.KLqprior().sum(list(range(-self.event_dim, 0))) + self.pi.KLqprior()

    def ELBO(self):
        return self.logZ - self.KLqprior()

    def assignment_pr(self):
        return self.p

    def assignment(self):
        return self.p.argmax(-1)

    def means(self):
        return self.dist.mean()

    def event_average_f(self, function_string, A=None, keepdim=False):
        f = getattr(self.dist, function_string)
        return self.event_average(f() if A is None else f(A), keepdim=keepdim)

    def average_f(self, function_string, A=None, keepdim=False):
        f = getattr(self.dist, function_string)
        return self.average(f() if A is None else f(A), keepdim=keepdim)

    def average(self, A, keepdim=False):
        return (A * self.p).sum(-1, keepdim)

    def event_average(self, A, keepdim=False):
        out = (A * self.p.view(self.p.shape + (1,) * self.dist.event_dim)).sum(-1 - self.dist.event_dim, keepdim)
        for i in range(self.event_dim - 1):
            out = out.sum(-self.dist.event_dim - 1, keepdim)
        return out


class GaussianMixtureModel(Mixture):
    def __init__(self, nc, dim, isotropic=False):
        """models/GaussianMixtureModel.py:7-12 (full-covariance branch; NormalGamma is SURVEY §8f #4)."""
        if isotropic is not False:
            raise NotImplementedError("isotropic=True (NormalGamma) is outside the NIW hot path (SURVEY.md §8f)")
        dist = NormalInverseWishart(event_shape=(dim,), batch_shape=(nc,), scale=1.0 / nc ** (1.0 / dim))
        super().__init__(dist, event_shape=(nc,))

    def initialize(self, data):
        """models/GaussianMixtureModel.py:14-16."""
        idx = torch.randint(data.shape[0], self.event_shape)
        self.dist.mu = data[idx, :]
