"""Gamma and DiagonalWishart nodes (dists/Gamma.py:6-111, dists/DiagonalWishart.py:7-66).  K x d floats per node:
element-wise torch on the node's device (the same call as the Dirichlet node); they carry the
precision state of the diagonal nodes NormalGamma / MatrixNormalGamma, whose O(N K d) passes are the kernels."""
from __future__ import annotations

import torch


class Gamma():
    def __init__(self, event_shape=(), batch_shape=(),
                 prior_parms={'alpha': torch.tensor(1.0, requires_grad=False),
                              'beta': torch.tensor(1.0, requires_grad=False)}):
        """dists/Gamma.py:7-23: alpha = alpha_0 + rand, beta = beta_0 + rand (two draws from the global RNG, in this order)."""
        self.event_dim = len(event_shape)
        self.event_shape = event_shape
        self.batch_dim = len(batch_shape)
        self.batch_shape = batch_shape
        self.nat_parms_0 = prior_parms
        dev = torch.empty(0).device
        self.alpha_0 = torch.as_tensor(prior_parms['alpha']).to(dev).expand(batch_shape + event_shape)
        self.beta_0 = torch.as_tensor(prior_parms['beta']).to(dev).expand(batch_shape + event_shape)
        self.alpha = self.alpha_0 + torch.rand(self.alpha_0.shape, requires_grad=False)
        self.beta = self.beta_0 + torch.rand(self.beta_0.shape, requires_grad=False)
        self.SEx = 0.0
        self.SElogx = 0.0

    def to_event(self, n):
        if n == 0:
            return self
        self.event_dim = self.event_dim + n
        self.batch_dim = self.batch_dim - n
        self.event_shape = self.batch_shape[-n:] + self.event_shape
        self.batch_shape = self.batch_shape[:-n]
        return self

    def to(self, device):
        for k in ("alpha_0", "beta_0", "alpha", "beta", "SEx", "SElogx"):
            if isinstance(getattr(self, k), torch.Tensor):
                setattr(self, k, getattr(self, k).to(device))
        return self

    def ss_update(self, SElogx, SEx, lr=1.0, beta=None):
        """dists/Gamma.py:34-47."""
        assert (SElogx.ndim == self.batch_dim + self.event_dim)
        assert (SEx.ndim == self.batch_dim + self.event_dim)
        if beta is not None:
            self.SEx = beta * self.SEx + SEx
            self.SElogx = beta * self.SElogx + SElogx
            SEx = self.SEx
            SElogx = self.SElogx
        self.alpha = (self.alpha_0 + SElogx) * lr + self.alpha * (1 - lr)
        self.beta = (self.beta_0 + SEx) * lr + self.beta * (1 - lr)

    def update(self, pX, p=None, lr=1.0, beta=None):
        """dists/Gamma.py:49-62: as raw_update, from a belief's mean."""
        self.raw_update(pX.mean(), p=p, lr=lr, beta=beta)

    def raw_update(self, X, p=None, lr=1.0, beta=None):
        """dists/Gamma.py:64-76 (Poisson observation model)."""
        sample_shape = X.shape[:-self.event_dim - self.batch_dim]
        sd = list(range(len(sample_shape)))
        if p is None:
            N = torch.tensor(float(torch.Size(sample_shape).numel()), device=X.device).expand(self.batch_shape + self.event_shape)
            SEx = X.sum(sd)
        else:
            p = p.view(p.shape + (1,) * self.event_dim)
            SEx = (X * p).sum(sd)
            N = p.sum(sd)
        self.ss_update(SEx, N, lr=lr, beta=beta)

    def Elog_like(self, X):
        return (X * self.loggeomean() - (X + 1).lgamma() - self.mean()).sum(list(range(-self.event_dim, 0)))

    def mean(self):
        return self.alpha / self.beta

    def var(self):
        return self.alpha / self.beta ** 2

    def meaninv(self):
        return self.beta / (self.alpha - 1)

    def ElogX(self):
        return self.alpha.digamma() - self.beta.log()

    def loggeomean(self):
        """dists/Gamma.py:102-103 (log alpha - log beta: the reference's definition, not the digamma one)."""
        return self.alpha.log() - self.beta.log()

    def entropy(self):
        return self.alpha.log() - self.beta.log() + self.alpha.lgamma() + (1 - self.alpha) * self.alpha.digamma()

    def logZ(self):
        return -self.alpha * self.beta.log() + self.alpha.lgamma()

    def logZprior(self):
        return -self.alpha_0 * self.beta_0.log() + self.alpha_0.lgamma()

    def KLqprior(self):
        """dists/Gamma.py:113-115."""
        KL = (self.alpha - self.alpha_0) * self.alpha.digamma() - self.alpha.lgamma() + self.alpha_0.lgamma() \
            + self.alpha_0 * (self.beta.log() - self.beta_0.log()) + self.alpha * (self.beta_0 / self.beta - 1)
        return KL.sum(list(range(-self.event_dim, 0)))


class DiagonalWishart():
    def __init__(self, event_shape, batch_shape=(), prior_parms={'nu': torch.tensor(2.0), 'U': torch.tensor(0.5)}, scale=1.0):
        """dists/DiagonalWishart.py:9-20: a Gamma node per diagonal element, alpha_0 = nu, beta_0 = scale^2 / U."""
        self.dim = event_shape[-1]
        self.event_dim = len(event_shape)
        self.event_shape = event_shape
        self.batch_dim = len(batch_shape)
        self.batch_shape = batch_shape
        self.gamma = Gamma(event_shape, batch_shape,
                           prior_parms={'alpha': prior_parms['nu'], 'beta': scale ** 2 / prior_parms['U']})

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
        self.gamma.to(device)
        return self

    def ss_update(self, SExx, N, lr=1.0, beta=None):
        """dists/DiagonalWishart.py:32-37: SExx is the DIAGONAL of a scatter matrix."""
        assert (SExx.ndim == self.batch_dim + self.event_dim)
        assert (N.ndim == self.batch_dim + self.event_dim)
        self.gamma.ss_update(N / 2.0, SExx / 2.0, lr, beta)

    def KLqprior(self):
        return self.gamma.KLqprior()

    def logZ(self):
        return self.gamma.logZ()

    def ESigma(self):
        return self.tensor_diag(self.gamma.meaninv())

    def EinvSigma(self):
        return self.tensor_diag(self.gamma.mean())

    def ElogdetinvSigma(self):
        return self.gamma.loggeomean().sum(-1)

    def logdetEinvSigma(self):
        return self.gamma.mean().log().sum(-1)

    def mean(self):
        return self.tensor_diag(self.gamma.mean())

    def tensor_diag(self, A):
        return A.unsqueeze(-1) * torch.eye(A.shape[-1], requires_grad=False, device=A.device)

    def tensor_extract_diag(self, A):
        return A.diagonal(dim1=-2, dim2=-1)
