"""NormalInverseWishart node with the reference's interface (dists/NormalInverseWishart.py:4-132).

Same constructor signature, attribute names and RNG consumption as the reference class, so it can be
bound in its place (see install.py); Elog_like / raw_update / ss_update / KLqprior run in
libvbmp_b200.so.  State stays in ordinary torch tensors that callers may read and overwrite.
"""
from __future__ import annotations

import math

import torch

from . import _lib, _shapes
from .wishart import Wishart


class NormalInverseWishart():

    def __init__(self, event_shape, batch_shape=(), scale=torch.tensor(1.0, requires_grad=False),
                 fixed_precision=False,
                 prior_parms={'lambda_mu': torch.tensor(1.0, requires_grad=False),
                              'mu': torch.tensor(0.0, requires_grad=False),
                              'nu': None,
                              'invU': None}):
        """dists/NormalInverseWishart.py:6-37 (same defaults; mu = mu_0 + randn consumes the global RNG)."""
        self.dim = event_shape[-1]
        self.event_shape = event_shape
        self.event_dim = len(event_shape)
        self.batch_shape = batch_shape
        self.batch_dim = len(batch_shape)
        self.fixed_precision = fixed_precision

        dev = torch.empty(0).device
        self.lambda_mu_0 = prior_parms['lambda_mu'].to(dev).expand(self.batch_shape + (self.event_dim - 1) * (1,))
        self.lambda_mu = self.lambda_mu_0
        self.mu_0 = prior_parms['mu'].to(dev).expand(self.batch_shape + event_shape)
        self.mu = self.mu_0 + torch.randn_like(self.mu_0, requires_grad=False)

        self.invU = Wishart(event_shape=event_shape + (self.dim,), batch_shape=batch_shape, scale=scale)
        if prior_parms['invU'] is not None and prior_parms['nu'] is not None:
            if self.invU.invU_0.shape == prior_parms['invU'].shape:
                self.invU.invU_0 = prior_parms['invU']
            else:
                print('Warning: NormalInverseWishart prior invU shape does not match Wishart invU_0 shape.  Using default.')
            if self.invU.nu_0.shape == prior_parms['nu'].shape:
                self.invU.nu_0 = prior_parms['nu']
            else:
                print('Warning: NormalInverseWishart prior nu shape does not match Wishart nu_0 shape.  Using default.')

        self.SExx = torch.tensor(0.0)
        self.SEx = torch.tensor(0.0)
        self.N = torch.tensor(0.0)

    def to_event(self, n):
        """dists/NormalInverseWishart.py:39-47."""
        if n == 0:
            return self
        self.event_dim = self.event_dim + n
        self.batch_dim = self.batch_dim - n
        self.event_shape = self.batch_shape[-n:] + self.event_shape
        self.batch_shape = self.batch_shape[:-n]
        self.invU.to_event(n)
        return self

    def to(self, device):
        for k in ("lambda_mu_0", "lambda_mu", "mu_0", "mu", "SExx", "SEx", "N"):
            setattr(self, k, getattr(self, k).to(device))
        self.invU.to(device)
        return self

    # ---- helpers ------------------------------------------------------------------------------------
    def _full(self):
        """Shapes of the per-component state: batch + extra-event dims, C of them."""
        full = tuple(self.batch_shape) + tuple(self.event_shape[:-1])
        return full, int(math.prod(full))

    def _state_flat(self):
        """Contiguous fp32 (C, ...) views of priors and posterior, in reference (batch-major) order."""
        full, C = self._full()
        d, dev = self.dim, self.mu.device
        f = _lib.f32
        w = self.invU
        return dict(
            C=C, d=d,
            lam0=f(self.lambda_mu_0.expand(full), dev).reshape(C), lam=f(self.lambda_mu.expand(full), dev).reshape(C),
            mu0=f(self.mu_0.expand(full + (d,)), dev).reshape(C, d), mu=f(self.mu.expand(full + (d,)), dev).reshape(C, d),
            invU0=f(w.invU_0.expand(full + (d, d)), dev).reshape(C, d, d), invU=f(w.invU.expand(full + (d, d)), dev).reshape(C, d, d),
            U=f(w.U.expand(full + (d, d)), dev).reshape(C, d, d),
            nu0=f(w.nu_0.expand(full), dev).reshape(C), nu=f(w.nu.expand(full), dev).reshape(C),
            logdet=f(w.logdet_invU.expand(full), dev).reshape(C), logdet0=f(w.logdet_invU_0.expand(full), dev).reshape(C),
        )

    def _plan(self, X):
        nb, ne = self.batch_dim, self.event_dim
        sample_shape = tuple(X.shape[:X.ndim - nb - ne])
        bstar = tuple(X.shape[X.ndim - nb - ne:X.ndim - ne])
        return _shapes.make_plan(self.batch_shape, self.event_shape[:-1], bstar, sample_shape)

    def _prep(self, plan, logprior=None):
        """K1: whitening factors from the live attributes, ordered (G, K)."""
        nb, nx, d, dev = self.batch_dim, self.event_dim - 1, self.dim, self.mu.device
        full, C = self._full()
        f = _lib.f32
        w = self.invU
        tk = lambda t, tail: _shapes.theta_to_GK(t, plan, nb, nx, tail)   # noqa: E731
        invU = tk(f(w.invU.expand(full + (d, d)), dev), 2)
        mu = tk(f(self.mu.expand(full + (d,)), dev), 1)
        nu = tk(f(w.nu.expand(full), dev), 0)
        lam = tk(f(self.lambda_mu.expand(full), dev), 0)
        lp = None
        if logprior is not None:
            lp = tk(f(logprior.expand(full), dev), 0)
        Dp = _lib.pad_dim(d)
        return _lib.niw_prep(invU, mu, nu, lam, lp, C, d, Dp) + (Dp,)

    # ---- reference protocol ---------------------------------------------------------------------------
    def ss_update(self, SExx, SEx, N, lr=1.0, beta=0.0):
        """dists/NormalInverseWishart.py:49-68 -> vbmp_niw_update (Wishart part included)."""
        assert (SExx.ndim == self.batch_dim + self.event_dim + 1)
        assert (SEx.ndim == self.batch_dim + self.event_dim)
        assert (N.ndim == self.batch_dim + self.event_dim - 1)

        if beta is not None:
            self.SExx = beta * self.SExx + SExx
            self.SEx = beta * self.SEx + SEx
            self.N = beta * self.N + N
            SExx = self.SExx
            SEx = self.SEx
            N = self.N
        s = self._state_flat()
        full, C = self._full()
        d, dev = self.dim, self.mu.device
        f = _lib.f32
        lam_shape = torch.broadcast_shapes(self.lambda_mu_0.shape, N.shape)
        lam, mu, invU, nu, U, logdet, info = _lib.niw_update(
            f(SExx.expand(full + (d, d)), dev).reshape(C, d, d), f(SEx.expand(full + (d,)), dev).reshape(C, d),
            f(torch.as_tensor(N, dtype=torch.float32, device=dev).expand(full), dev).reshape(C),
            s["lam0"], s["mu0"], s["invU0"], s["nu0"], s["lam"], s["mu"], s["invU"], s["nu"],
            C, d, float(lr), self.fixed_precision is not False)
        lam = lam.view(full)
        # keep the reference's broadcast shape of lambda_mu (e.g. (K,1) when extra event dims share N)
        idx = tuple(slice(0, 1) if ls == 1 else slice(None) for ls in lam_shape)
        self.lambda_mu = lam[idx]
        self.mu = mu.view(full + (d,))
        if self.fixed_precision is False:
            self.invU._set(invU, nu, U, logdet, info)

    def _gram(self, X, p=None):
        """K3: weighted Gram statistics of z = [x;1] in kernel layout (G, K, d+1, d+1)."""
        plan = self._plan(X)
        dev = self.mu.device
        d = self.dim
        Xc = _lib.f32(X, dev).reshape(plan.N, plan.GX, d)
        pc = None
        if p is not None:
            pc = _lib.f32(p, dev).reshape(plan.N, plan.GP, plan.K)
        xg = _shapes.idx_tensor(plan.xg, dev)
        pg = _shapes.idx_tensor(plan.pg, dev)
        return _lib.gram(Xc, None, plan.N, plan.GX, xg, pc, plan.GP, pg, plan.G, plan.K, _lib.pad_dim(d)), plan

    def _update_from_gram(self, G, plan, weighted, lr=1.0, beta=None, n_samples=None):
        """Blocks of the Gram matrix are SExx / SEx / N of dists/NormalInverseWishart.py:80-84."""
        d, dev = self.dim, self.mu.device
        G = _shapes.GK_to_theta(G, plan, (d + 1, d + 1))          # batch + extra + (d+1, d+1)
        SExx = G[..., :d, :d]
        SEx = G[..., :d, d]
        if not weighted:
            n = float(plan.N if n_samples is None else n_samples)
            N = torch.tensor(n, device=dev).expand(self.batch_shape + self.event_shape[:-1])
        else:
            # the reference's N is p summed over samples, viewed with singleton extra-event dims (:80-81)
            N = G[..., d, d][(Ellipsis,) + (slice(0, 1),) * (self.event_dim - 1)]
        self.ss_update(SExx, SEx, N, lr, beta)

    def raw_update(self, X, p=None, lr=1.0, beta=None):
        """dists/NormalInverseWishart.py:70-86: one weighted Gram pass (vbmp_gram, z = [x;1]) feeds ss_update."""
        G, plan = self._gram(X, p)
        self._update_from_gram(G, plan, p is not None, lr, beta)

    def update(self, pX, p=None, lr=1.0, beta=None):
        """dists/NormalInverseWishart.py:88-89 (a stub in the reference as well)."""
        pass

    def Elog_like(self, X):
        """dists/NormalInverseWishart.py:91-97 -> K1 + K2 (mode 0: logits only)."""
        plan = self._plan(X)
        dev = self.mu.device
        W, m, cst, info, Dp = self._prep(plan)
        Xc = _lib.f32(X, dev).reshape(plan.N, plan.GX, self.dim)
        out = _lib.estep(Xc, None, plan.N, plan.GX, _shapes.idx_tensor(plan.xg, dev), W, m, cst,
                         plan.G, plan.K, Dp, 0)
        out = _shapes.logits_to_ref(out, plan)
        for i in range(self.event_dim - 1):
            out = out.sum(-1)
        return out

    def KLqprior(self):
        """dists/NormalInverseWishart.py:99-105 -> vbmp_niw_kl."""
        s = self._state_flat()
        full, C = self._full()
        KL = _lib.niw_kl(s["lam0"], s["lam"], s["mu0"], s["mu"], s["invU0"], s["U"], s["nu0"], s["nu"],
                         s["logdet"], s["logdet0"], C, self.dim).view(full)
        for i in range(self.event_dim - 1):
            KL = KL.sum(-1)
        return KL

    def mean(self):
        return self.mu

    def EX(self):
        return self.mu

    def EXXT(self):
        return self.mu.unsqueeze(-1) * self.mu.unsqueeze(-2) + self.invU.ESigma() / self.lambda_mu.unsqueeze(-1).unsqueeze(-1)

    def ESigma(self):
        return self.invU.ESigma()

    def ElogdetinvSigma(self):
        return self.invU.ElogdetinvSigma()

    def EinvSigmamu(self):
        return (self.invU.EinvSigma() * self.mu.unsqueeze(-2)).sum(-1)

    def EinvSigma(self):
        return self.invU.EinvSigma()

    def EinvUX(self):
        return (self.invU.EinvSigma() * self.mu.unsqueeze(-2)).sum(-1)

    def EXTinvUX(self):
        return (self.mu.unsqueeze(-1) * self.invU.EinvSigma() * self.mu.unsqueeze(-2)).sum(-1).sum(-1) \
            + self.dim / self.lambda_mu
