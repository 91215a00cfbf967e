"""MultivariateNormal_vector_format with the reference's interface (dists/MultivariateNormal_vector_format.py:3-119):
the value type `predict` returns.  A plain container of moments / natural parameters with lazy conversions; nothing here
is on the VB-EM hot path (batches of n x n inverses of an (N, n, n) result are torch plumbing)."""
from __future__ import annotations

import torch


class MultivariateNormal_vector_format():

    def __init__(self, mu=None, Sigma=None, invSigmamu=None, invSigma=None, logdetinvSigma=None):
        """dists/MultivariateNormal_vector_format.py:4-27: vectors are (dim, 1) matrices."""
        self.mu = mu
        self.Sigma = Sigma
        self.invSigmamu = invSigmamu
        self.invSigma = invSigma
        self.logdetinvSigma = logdetinvSigma
        self.event_dim = 2
        if self.mu is not None:
            self.dim = mu.shape[-2]
            self.event_shape = mu.shape[-2:]
            self.batch_shape = mu.shape[:-2]
        elif self.invSigmamu is not None:
            self.dim = invSigmamu.shape[-2]
            self.event_shape = invSigmamu.shape[-2:]
            self.batch_shape = invSigmamu.shape[:-2]
        else:
            print('mu and invSigmamu are both None: cannont initialize MultivariateNormal')
            return None
        self.batch_dim = len(self.batch_shape)
        self.event_dim = len(self.event_shape)

    @property
    def shape(self):
        return self.batch_shape + self.event_shape

    def mean(self):                                                     # :79-82
        if self.mu is None:
            self.mu = self.invSigma.inverse() @ self.invSigmamu
        return self.mu

    def ESigma(self):                                                   # :84-87
        if self.Sigma is None:
            self.Sigma = self.invSigma.inverse()
        return self.Sigma

    def EinvSigma(self):                                                # :89-92
        if self.invSigma is None:
            self.invSigma = self.Sigma.inverse()
        return self.invSigma

    def EinvSigmamu(self):                                              # :94-97
        if self.invSigmamu is None:
            self.invSigmamu = self.EinvSigma() @ self.mean()
        return self.invSigmamu

    def ElogdetinvSigma(self):                                          # :104-107
        if self.logdetinvSigma is None:
            self.logdetinvSigma = self.EinvSigma().logdet()
        return self.logdetinvSigma

    def EX(self):
        return self.mean()

    def EXXT(self):                                                     # :112-113
        return self.ESigma() + self.mean() @ self.mean().transpose(-2, -1)

    def EXTX(self):                                                     # :115-116
        return self.ESigma().sum(-1).sum(-1) + (self.mean().transpose(-2, -1) @ self.mean()).squeeze(-1).squeeze(-1)

    def Res(self):                                                      # :118-119
        return (- 0.5 * (self.mean() * self.EinvSigmamu()).sum(-1).sum(-1) + 0.5 * self.ElogdetinvSigma()
                - 0.5 * self.dim * torch.log(2 * torch.tensor(torch.pi, requires_grad=False)))
