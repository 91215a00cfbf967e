"""Dirichlet node (dists/Dirichlet.py:3-86).  K floats per mixture: stays in torch on the node's
device ("tiny, K floats; stays in torch"); it only feeds log-prior terms to K1."""
from __future__ import annotations

import torch


class Dirichlet():
    def __init__(self, event_shape, batch_shape=(), prior_parms={'alpha': torch.tensor(0.5)}):
        """dists/Dirichlet.py:4-11: alpha = alpha_0 (1 + rand) consumes the global RNG."""
        self.event_dim = len(event_shape)
        self.batch_dim = len(batch_shape)
        self.event_shape = event_shape
        self.batch_shape = batch_shape
        dev = torch.empty(0).device
        self.alpha_0 = prior_parms['alpha'].to(dev).expand(batch_shape + event_shape)
        self.alpha = self.alpha_0 * (1.0 + torch.rand(self.alpha_0.shape, requires_grad=False))
        self.NA = 0.0

    def to_event(self, n):
        if n == 0:
            return self
        self.event_dim = self.event_dim + n
        self.batch_dim = self.batch_dim - n
        self.event_shape = self.batch_shape[-n:] + self.event_shape
        self.batch_shape = self.batch_shape[:-n]
        return self

    def to(self, device):
        self.alpha_0 = self.alpha_0.to(device)
        self.alpha = self.alpha.to(device)
        if isinstance(self.NA, torch.Tensor):
            self.NA = self.NA.to(device)
        return self

    def _ed(self):
        return list(range(-self.event_dim, 0))

    def ss_update(self, NA, lr=1.0, beta=None):
        """dists/Dirichlet.py:22-28."""
        assert (NA.shape == self.batch_shape + self.event_shape)
        self.NA = beta * self.NA + NA if beta is not None else NA
        self.alpha = lr * (self.NA + self.alpha_0) + (1 - lr) * self.alpha

    def raw_update(self, X, p=None, lr=1.0, beta=None):
        """dists/Dirichlet.py:30-37."""
        sample_dim = X.ndim - self.event_dim - self.batch_dim
        if p is None:
            NA = X.sum(list(range(sample_dim)))
        else:
            NA = (X * p.view(p.shape + (1,) * self.event_dim)).sum(list(range(sample_dim)))
        self.ss_update(NA, lr, beta)

    def update(self, X, p=None, lr=1.0, beta=None):
        self.raw_update(X, p, lr, beta)

    def mean(self):
        return self.alpha / self.alpha.sum(self._ed(), keepdim=True)

    def loggeomean(self):
        """dists/Dirichlet.py:52-53."""
        return self.alpha.digamma() - self.alpha.sum(self._ed(), keepdim=True).digamma()

    def ElogX(self):
        return self.loggeomean()

    def var(self):
        s = self.alpha.sum(self._ed(), keepdim=True)
        mean = self.mean()
        return mean * (1 - mean) / (s + 1)

    def Elog_like(self, X):
        ed = self._ed()
        return (X * self.loggeomean()).sum(ed) + (1 + X.sum(ed)).lgamma() - (1 + X).lgamma().sum(ed)

    def KL_lgamma(self, x):
        """dists/Dirichlet.py:63-66 (infinite entries contribute zero; not in place here)."""
        return torch.nan_to_num(x.lgamma(), posinf=0.0)

    def KL_digamma(self, x):
        """dists/Dirichlet.py:68-71."""
        return torch.nan_to_num(x.digamma(), neginf=0.0)

    def KLqprior(self):
        """dists/Dirichlet.py:73-83 (infinite lgamma / digamma entries contribute zero)."""
        ed = self._ed()
        a, a0 = self.alpha, self.alpha_0
        lg = lambda x: torch.nan_to_num(x.lgamma(), posinf=0.0)               # noqa: E731
        dg = lambda x: torch.nan_to_num(x.digamma(), neginf=0.0)              # noqa: E731
        asum = a.sum(ed)
        KL = asum.lgamma() - lg(a).sum(ed) - a0.sum(ed).lgamma() + lg(a0).sum(ed)
        KL = KL + ((a - a0) * (dg(a) - asum.digamma().view(asum.shape + (1,) * self.event_dim))).sum(ed)
        while KL.ndim > self.batch_dim:
            KL = KL.sum(-1)
        return KL

    def logZ(self):
        ed = self._ed()
        return self.alpha.lgamma().sum(ed) - self.alpha.sum(ed).lgamma()
