"""HMM / ARHMM glue with the reference's interface (models/HMM.py:5-178, models/ARHMM.py:13-25).

The emission E-step (obs_logits -> K1 + K2, mode 0) and M-step (raw_update -> weighted Gram +
update kernel) run in libvbmp_b200.so with the (T, S) sample axes flattened.  The forward-backward
recursion between them is one CUDA kernel for K <= 32 states (csrc/hmm_fb.cu).
"""
from __future__ import annotations

import torch

from .dirichlet import Dirichlet
from .mnw import MatrixNormalWishart


def _lse(x, dim, keepdim=False):
    return torch.logsumexp(x, dim=dim, keepdim=keepdim)


class HMM():
    def __init__(self, obs_dist, transition_mask=None, ptemp=1.0):
        """models/HMM.py:6-31."""
        self.obs_dist = obs_dist
        self.event_dim = 1
        self.dim = obs_dist.batch_shape[-1]
        self.event_shape = obs_dist.batch_shape[-1:]
        self.batch_shape = obs_dist.batch_shape[:-1]
        self.batch_dim = len(self.batch_shape)
        self.transition_mask = transition_mask

        alpha = torch.eye(self.dim, requires_grad=False) + 0.5
        if transition_mask is not None:
            alpha = alpha * transition_mask
        self.transition = Dirichlet(self.event_shape, self.batch_shape + self.event_shape, prior_parms={'alpha': alpha})
        self.initial = Dirichlet(self.event_shape, self.batch_shape)

        self.sumlogZ = -torch.inf
        self.p = None
        self.ptemp = ptemp
        self.logZ = torch.tensor(-torch.inf)
        self.ELBO_last = torch.tensor(-torch.inf)

    def to(self, device):
        self.obs_dist.to(device)
        self.transition.to(device)
        self.initial.to(device)
        self.logZ = self.logZ.to(device)
        self.ELBO_last = self.ELBO_last.to(device)
        return self

    def forward_backward_logits(self, fw_logits):
        """models/HMM.py:72-105: log-space filter, backward smoother, expected transition counts.

        On the device with K <= 32 states this is ONE kernel (vbmp_hmm_forward_backward: a warp per sequence for the
        whole recursion); otherwise the same recursion as batched torch ops on the tensor's device."""
        tr = self.transition.loggeomean()
        init = self.initial.loggeomean()
        T = fw_logits.shape[0]
        K = fw_logits.shape[-1]
        if fw_logits.is_cuda and K <= 32 and T >= 1 and fw_logits.numel() > 0:
            from . import _lib
            dev = fw_logits.device
            rest = tuple(fw_logits.shape[1:-1])                  # other sample dims + batch dims (batch dims last)
            G = 1
            for b in self.batch_shape:
                G *= int(b)
            S = 1
            for r in rest:
                S *= int(r)
            lg = _lib.f32(fw_logits, dev).reshape(T, S, K)
            trf = _lib.f32(tr.expand(tuple(self.batch_shape) + (K, K)), dev).reshape(G, K, K)
            inf = _lib.f32(init.expand(tuple(self.batch_shape) + (K,)), dev).reshape(G, K)
            p, SEzz, SEz0, logZ = _lib.hmm_forward_backward(lg, trf, inf, T, S, G, K, float(self.ptemp))
            return (p.view(fw_logits.shape), SEzz.view(rest + (K, K)), SEz0.view(rest + (K,)), logZ.view(rest))
        return self._forward_backward_torch(fw_logits, tr, init)

    def _forward_backward_torch(self, fw_logits, tr, init):
        T = fw_logits.shape[0]
        fw = torch.empty_like(fw_logits)
        fw[0] = _lse(init.unsqueeze(-1) + tr + fw_logits[0].unsqueeze(-2), -2)
        for t in range(1, T):
            fw[t] = _lse(fw[t - 1].unsqueeze(-1) + tr + fw_logits[t].unsqueeze(-2), -2)
        logZ = _lse(fw[-1], -1, True)
        fw = fw - logZ
        logZ = logZ.squeeze(-1)
        SEzz = torch.zeros(fw.shape[1:] + self.event_shape, dtype=fw.dtype, device=fw.device)
        for t in range(T - 2, -1, -1):
            temp = fw[t].unsqueeze(-1) + tr
            xi = (temp - _lse(temp, -2, True)) + fw[t + 1].unsqueeze(-2)
            fw[t] = _lse(xi, -1)
            SEzz = SEzz + (xi - _lse(xi, (-1, -2), True)).exp()
        temp = init.unsqueeze(-1) + tr
        xi = (temp - _lse(temp, -2, True)) + fw[0].unsqueeze(-2)
        SEz0 = _lse(xi, -1)
        SEz0 = (SEz0 - _lse(SEz0, -1, True)).exp()
        SEzz = SEzz + (xi - _lse(xi, (-1, -2), True)).exp()
        p = ((fw - fw.max(-1, keepdim=True)[0]) / self.ptemp).exp()
        p = p / p.sum(-1, keepdim=True)
        return p, SEzz, SEz0, logZ

    def forward_step(self, logits, observation_logits):
        """models/HMM.py:33-34."""
        return _lse(logits.unsqueeze(-1) + observation_logits.unsqueeze(-2) + self.transition.loggeomean(), -2)

    def backward_step(self, logits, observation_logits):
        """models/HMM.py:36-37."""
        return _lse(logits.unsqueeze(-2) + observation_logits.unsqueeze(-2) + self.transition.loggeomean(), -1)

    def forward_backward_steps(self, X, T):
        """models/HMM.py:39-71: the same recursion with the observation logits evaluated one time slice at a time."""
        return self.forward_backward_logits(torch.stack([self.obs_logits(X, t) for t in range(T)], 0))

    def assignment_pr(self):
        return self.p

    def assignment(self):
        return self.p.argmax(-1)

    def obs_logits(self, X, t=None):
        """models/HMM.py:113-117."""
        if t is not None:
            return self.obs_dist.Elog_like(X[t].unsqueeze(-1 - self.obs_dist.event_dim))
        return self.obs_dist.Elog_like(X.unsqueeze(-1 - self.obs_dist.event_dim))

    def update_states(self, X, T=None):
        """models/HMM.py:119-132 (T=None path)."""
        if T is not None:          # models/HMM.py:120-121 (the reference's own version stops at an undefined helper, :61)
            self.p, SEzz, SEz0, logZ = self.forward_backward_steps(X, T)
        else:
            self.p, SEzz, SEz0, logZ = self.forward_backward_logits(self.obs_logits(X))
        NA = self.p.sum(0)
        sample_dims = list(range(NA.ndim - self.batch_dim - self.event_dim))
        NA = NA.sum(sample_dims)
        SEzz = SEzz.sum(sample_dims)
        SEz0 = SEz0.sum(sample_dims)
        logZ = logZ.sum(sample_dims)
        return SEzz, SEz0, NA, logZ

    def update_markov_parms(self, SEzz, SEz0, lr=1.0, beta=None):
        self.transition.ss_update(SEzz, lr=lr, beta=beta)
        self.initial.ss_update(SEz0, lr=lr, beta=beta)

    def update_obs_parms(self, X, lr=1.0, beta=None):
        """models/HMM.py:138-139."""
        self.obs_dist.raw_update(X.unsqueeze(-1 - self.obs_dist.event_dim), p=self.p, lr=lr, beta=beta)

    def update(self, X, iters=1, T=None, lr=1.0, beta=None, verbose=False):
        """models/HMM.py:141-152 (ELBO is evaluated after the M-step here)."""
        for i in range(iters):
            SEzz, SEz0, self.NA, self.logZ = self.update_states(X, T)
            self.KLqprior_last = self.KLqprior()
            self.update_markov_parms(SEzz, SEz0, lr=lr, beta=beta)
            self.update_obs_parms(X, lr=lr, beta=beta)
            ELBO = self.ELBO()
            if verbose:
                print('Percent Change in ELBO = ', ((ELBO - self.ELBO_last) / torch.abs(self.ELBO_last) * 100))
            self.ELBO_last = ELBO

    def KLqprior(self):
        return self.obs_dist.KLqprior().sum(-1) + self.transition.KLqprior().sum(-1) + self.initial.KLqprior()

    def ELBO(self):
        return self.logZ - self.KLqprior()

    def average(self, A, keepdim=False):
        return (A * self.p).sum(-1, keepdim)

    def event_average(self, A, keepdim=False):
        out = (A * self.p.view(self.p.shape + (1,) * self.obs_dist.event_dim)).sum(-self.obs_dist.event_dim - 1, keepdim)
        for i in range(self.event_dim - 1):
            out = out.sum(-self.obs_dist.event_dim - 1, keepdim)
        return out

    def event_average_f(self, function_string, keepdim=False):
        return self.event_average(getattr(self.obs_dist, function_string)(), keepdim)

    def average_f(self, function_string, keepdim=False):
        return self.average(getattr(self.obs_dist, function_string)(), keepdim)


class ARHMM(HMM):
    def __init__(self, dim, n, p, batch_shape=(), pad_X=True, X_mask=None, mask=None, transition_mask=None):
        """models/ARHMM.py:14-16."""
        dist = MatrixNormalWishart(event_shape=(n, p), batch_shape=batch_shape + (dim,), pad_X=pad_X,
                                   X_mask=X_mask, mask=mask)
        super().__init__(dist, transition_mask=transition_mask)

    def obs_logits(self, XY, t=None):
        """models/ARHMM.py:18-22."""
        if t is not None:
            return self.obs_dist.Elog_like(XY[0][t], XY[1][t])
        return self.obs_dist.Elog_like(XY[0], XY[1])

    def update_obs_parms(self, XY, lr, beta):
        """models/ARHMM.py:24-25."""
        self.obs_dist.raw_update(XY[0], XY[1], p=self.p, lr=lr, beta=beta)


class ARHMM_prXY(HMM):
    """models/ARHMM.py:35-46: the ARHMM driven by Gaussian beliefs (pX, pY) about regressors and outputs instead of raw data
    (SURVEY.md §8f #2).  Observation logits are the expected log likelihoods, the observation update uses expected
    sufficient statistics; both run on the E-step / Gram kernels over the means plus covariance corrections."""

    def __init__(self, dim, n, p, batch_shape=(), X_mask=None, mask=None, pad_X=True, transition_mask=None):
        dist = MatrixNormalWishart(event_shape=(n, p), batch_shape=batch_shape + (dim,), pad_X=pad_X,
                                   X_mask=X_mask, mask=mask)
        super().__init__(dist, transition_mask=transition_mask)

    def obs_logits(self, XY, t=None):
        if t is not None:
            raise NotImplementedError("time-sliced beliefs (HMM.update with T) are not supported for ARHMM_prXY")
        return self.obs_dist.Elog_like_given_pX_pY(XY[0], XY[1])

    def update_obs_parms(self, XY, lr, beta=None):
        self.obs_dist.update(XY[0], XY[1], self.p, lr=lr, beta=beta)

    def Elog_like_X_given_pY(self, pY):
        raise NotImplementedError("message passing to X is outside the VB-EM hot path (SURVEY.md §2.1 #5)")


class ARHMM_prXRY(HMM):
    """models/ARHMM.py:55-91: the ARHMM whose regressors are a belief about latent X (a MultivariateNormal_vector_format)
    stacked on OBSERVED regressors R, with observed outputs Y — DynamicMarkovBlanketDiscovery's observation model.  The
    stacked regressor belief has mean [E x; r] and the block-diagonal covariance diag(Sigma_x, 0), so the observation
    logits and the observation update are MatrixNormalWishart.Elog_like_given_pX_pY / update on the kernels."""

    def __init__(self, dim, n, p1, p2, batch_shape=(), mask=None, X_mask=None, transition_mask=None, pad_X=False):
        self.p1 = p1
        self.p2 = p2
        dist = MatrixNormalWishart(event_shape=(n, p1 + p2), batch_shape=batch_shape + (dim,), pad_X=pad_X,
                                   X_mask=X_mask, mask=mask)
        super().__init__(dist, transition_mask=transition_mask)

    def _stack(self, XRY):
        from .mvn import MultivariateNormal_vector_format
        Sx = XRY[0].ESigma()
        lead, p1, p2 = Sx.shape[:-2], self.p1, self.p2
        Sigma = torch.zeros(lead + (p1 + p2, p1 + p2), dtype=Sx.dtype, device=Sx.device)
        Sigma[..., :p1, :p1] = Sx                                           # utils/matrix_utils.py:4-9 with a zero R block
        mu = torch.cat((XRY[0].mean(), XRY[1]), dim=-2)
        return MultivariateNormal_vector_format(mu=mu, Sigma=Sigma)

    def Elog_like(self, XRY):
        return (self.obs_logits(XRY) * self.p).sum(-1)

    def obs_logits(self, XRY, t=None):
        from .mvn import Delta
        if t is not None:
            raise NotImplementedError("time-sliced beliefs (HMM.update with T) are not supported for ARHMM_prXRY")
        return self.obs_dist.Elog_like_given_pX_pY(self._stack(XRY), Delta(XRY[2]))

    def update_obs_parms(self, XRY, lr, beta=None):
        from .mvn import Delta
        self.obs_dist.update(self._stack(XRY), Delta(XRY[2]), p=self.p, lr=lr, beta=beta)
