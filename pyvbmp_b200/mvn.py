"""Value type returned by `predict`: a (batch of) Gaussian over column vectors held either by moments (mu, Sigma) or by
natural parameters (invSigma, invSigmamu), with the accessor names of the reference's
dists/MultivariateNormal_vector_format.py (mean / ESigma / EinvSigma / EinvSigmamu / ElogdetinvSigma / EX / EXXT / EXTX /
Res).  Whichever representation is missing is derived on first use and cached.  Nothing here is on the VB-EM hot path;
the conversions are batched n x n inverses on the device."""
from __future__ import annotations

import math

import torch


class MultivariateNormal_vector_format():
    event_dim = 2            # vectors are (dim, 1) matrices

    def __init__(self, mu=None, Sigma=None, invSigmamu=None, invSigma=None, logdetinvSigma=None):
        self.mu, self.Sigma = mu, Sigma
        self.invSigmamu, self.invSigma = invSigmamu, invSigma
        self.logdetinvSigma = logdetinvSigma
        vec = mu if mu is not None else invSigmamu
        if vec is None:
            raise ValueError("MultivariateNormal_vector_format needs mu or invSigmamu")
        self.dim = vec.shape[-2]
        self.event_shape = tuple(vec.shape[-2:])
        self.batch_shape = tuple(vec.shape[:-2])
        self.batch_dim = len(self.batch_shape)

    @property
    def shape(self):
        return self.batch_shape + self.event_shape

    def unsqueeze(self, dim):
        """A new belief with a singleton batch axis at `dim` (counted from the end, in front of the (dim, 1) event)."""
        assert dim + self.event_dim < 0
        un = lambda t: None if t is None else t.unsqueeze(dim)      # noqa: E731
        return MultivariateNormal_vector_format(un(self.mu), un(self.Sigma), un(self.invSigmamu), un(self.invSigma))

    # ---- the two representations, each filled in from the other on demand ---------------------------------
    def ESigma(self):
        if self.Sigma is None:
            self.Sigma = torch.linalg.inv(self.invSigma)
        return self.Sigma

    def EinvSigma(self):
        if self.invSigma is None:
            self.invSigma = torch.linalg.inv(self.Sigma)
        return self.invSigma

    def mean(self):
        if self.mu is None:
            self.mu = self.ESigma() @ self.invSigmamu
        return self.mu

    def EinvSigmamu(self):
        if self.invSigmamu is None:
            self.invSigmamu = self.EinvSigma() @ self.mean()
        return self.invSigmamu

    def ElogdetinvSigma(self):
        if self.logdetinvSigma is None:
            self.logdetinvSigma = torch.logdet(self.EinvSigma())
        return self.logdetinvSigma

    # ---- expectations -----------------------------------------------------------------------------------
    EX = mean

    def EXXT(self):
        m = self.mean()
        return self.ESigma() + m @ m.transpose(-2, -1)

    def EXTX(self):
        m = self.mean()
        return self.ESigma().sum((-2, -1)) + (m * m).sum((-2, -1))

    def Res(self):
        """Log normaliser of the natural form: -mu^T invSigma mu / 2 + logdet(invSigma) / 2 - dim log(2 pi) / 2."""
        quad = (self.mean() * self.EinvSigmamu()).sum((-2, -1))
        return 0.5 * (self.ElogdetinvSigma() - quad - self.dim * math.log(2.0 * math.pi))


class Delta():
    """dists/Delta.py:6-52: an observation dressed as a distribution (a point mass at X), so that callers can ask it for the
    expectations a belief provides.  ESigma (not in the reference class) is the zero covariance, as a scalar that broadcasts."""

    def __init__(self, X):
        self.X = X

    def unsqueeze(self, dim):
        return Delta(self.X.unsqueeze(dim))

    def squeeze(self, dim):
        return Delta(self.X.squeeze(dim))

    def sum(self, dim, keepdim=False):
        return self.X.sum(dim, keepdim=keepdim)

    def cumsum(self, dim):
        return self.X.cumsum(dim)

    @property
    def shape(self):
        return self.X.shape

    def mean(self):
        return self.X

    def EX(self):
        return self.X

    def EXXT(self):
        return self.X @ self.X.transpose(-1, -2)

    def EXTX(self):
        return self.X.transpose(-1, -2) @ self.X

    def EXTAX(self, A):
        return self.X.transpose(-1, -2) @ A @ self.X

    def EXX(self):
        return self.X ** 2

    def ElogX(self):
        return torch.log(self.X)

    def E(self, f):
        return f(self.X)
