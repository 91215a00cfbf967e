"""MatrixNormalGamma node — MatrixNormalWishart with a DIAGONAL output precision (a Gamma node per output) — with the
reference's interface for the VB-EM hot path (transforms/MatrixNormalGamma.py:10-471).  It shares every
kernel with MatrixNormalWishart: K1 through vbmp_mnw_prep_ex (E[invSigma] = diag(alpha / beta), E log det = sum log alpha -
log beta), K2 (E-step) and K3 (weighted Gram over z = [x; y; 1]) unchanged, the invV / mu part of K5 (vbmp_mnw_update);
the Gamma part of the update and the KL are K x n element-wise torch on the device.  mask / X_mask and the
message-passing methods are outside the scope table: the class pyvbmp_b200.install() binds inherits them from the reference."""
from __future__ import annotations

import torch

from . import _lib, _shapes
from .gamma import DiagonalWishart
from .mnw import MatrixNormalWishart


class MatrixNormalGamma(MatrixNormalWishart):

    def __init__(self, event_shape, batch_shape=(), prior_parms={'mu': torch.tensor(0.0)}, scale=1.0,
                 uniform_precision=False, mask=None, X_mask=None, pad_X=False, fixed_precision=False):
        """transforms/MatrixNormalGamma.py:22-86 (mu = randn / sqrt(p'), then the Gamma node's two rand draws)."""
        if mask is not None or X_mask is not None:
            raise NotImplementedError("mask / X_mask are outside the accelerated path (SURVEY.md §2.1 #5)")
        self.n = event_shape[-2]
        self.p = event_shape[-1]
        self.pad_X = pad_X
        self.fixed_precision = fixed_precision
        self.uniform_precision = uniform_precision
        dev = torch.empty(0).device
        mu_0 = prior_parms['mu'].to(dev)
        if pad_X:
            self.p = self.p + 1
            event_shape = event_shape[:-1] + (self.p,)
            if mu_0.ndim != 0:
                mu_0 = torch.cat((mu_0, torch.zeros(mu_0.shape[:-1] + (1,), requires_grad=False)), dim=-1).clone()
        mu_0 = mu_0.expand(batch_shape + event_shape)
        self.event_dim = len(event_shape)
        self.event_shape = event_shape
        self.batch_dim = len(batch_shape)
        self.batch_shape = batch_shape
        self.mask = None
        self.X_mask = None
        self.mu_0 = mu_0
        self.mu = torch.randn_like(mu_0, requires_grad=False) / torch.sqrt(torch.tensor(float(self.p), requires_grad=False))
        mshape = batch_shape + event_shape[:-2]
        self.invV_0 = torch.eye(self.p, requires_grad=False).expand(mshape + (self.p, self.p))
        self.invV = self.invV_0
        self.V = self.invV_0
        self.invU = DiagonalWishart(event_shape=event_shape[:-1], batch_shape=batch_shape, scale=scale)
        self.logdetinvV = torch.zeros(mshape)
        self.logdetinvV_0 = torch.zeros(mshape)
        self.SEyy = 0.0
        self.SExx = 0.0
        self.SEyx = 0.0
        self.N = 0.0
        self.log2pi = torch.tensor(2 * torch.pi, requires_grad=False).log()

    def _prep(self, plan, logprior=None):
        nb, nx, dev = self.batch_dim, self.event_dim - 2, self.mu.device
        full, C = self._full()
        n, pp = self.n, self.p
        f = _lib.f32
        tk = lambda t, tail: _shapes.theta_to_GK(t, plan, nb, nx, tail)   # noqa: E731
        g = self.invU.gamma
        tau = tk(f(g.mean().expand(full + (n,)), dev), 1)
        eld = tk(f(g.loggeomean().sum(-1).expand(full), dev), 0)
        mu = tk(f(self.mu.expand(full + (n, pp)), dev), 2)
        invV = tk(f(self.invV.expand(full + (pp, pp)), dev), 2)
        lp = None if logprior is None else tk(f(logprior.expand(full), dev), 0)
        D = n + pp - int(self.pad_X)
        Dp = _lib.pad_dim(D)
        return _lib.mnw_prep(None, None, mu, invV, lp, C, n, pp, self.pad_X, Dp, tau=tau, elogdet=eld) + (Dp,)

    def _state_flat(self):
        full, C = self._full()
        n, pp, dev = self.n, self.p, self.mu.device
        f = _lib.f32
        return dict(
            C=C,
            mu0=f(self.mu_0.expand(full + (n, pp)), dev).reshape(C, n, pp), mu=f(self.mu.expand(full + (n, pp)), dev).reshape(C, n, pp),
            invV0=f(self.invV_0.expand(full + (pp, pp)), dev).reshape(C, pp, pp), invV=f(self.invV.expand(full + (pp, pp)), dev).reshape(C, pp, pp),
        )

    def ss_update(self, SExx, SEyx, SEyy, N, lr=1.0, beta=None):
        """transforms/MatrixNormalGamma.py:87-141, no-mask branch: invV / mu through vbmp_mnw_update (its fixed-precision
        path), the Gamma node from the DIAGONAL of SEyy - mu invV mu^T + mu_0 invV_0 mu_0^T (:123-124, un-blended mu, invV)."""
        assert (SExx.ndim == self.batch_dim + self.event_dim)
        assert (SEyx.ndim == self.batch_dim + self.event_dim)
        assert (SEyy.ndim == self.batch_dim + self.event_dim)
        assert (N.ndim == self.batch_dim + self.event_dim - 2)
        if beta is not None:
            self.SExx = beta * self.SExx + SExx
            self.SEyx = beta * self.SEyx + SEyx
            self.SEyy = beta * self.SEyy + SEyy
            self.N = beta * self.N + N
            SExx, SEyx, SEyy, N = self.SExx, self.SEyx, self.SEyy, self.N
        s = self._state_flat()
        full, C = self._full()
        n, pp, dev = self.n, self.p, self.mu.device
        f = _lib.f32
        sxx = f(SExx.expand(full + (pp, pp)), dev).reshape(C, pp, pp)
        syx = f(SEyx.expand(full + (n, pp)), dev).reshape(C, n, pp)
        Nf = f(torch.as_tensor(N, dtype=torch.float32, device=dev).expand(full), dev).reshape(C)

        def solve(lr_):
            out = _lib.mnw_update(sxx, syx, None, Nf, s["mu0"], s["invV0"], None, None, s["mu"], s["invV"], None, None,
                                  C, n, pp, float(lr_), True)
            return out[0], out[1], out[2], out[3], out[8]
        mu1, invV1, V1, ld1, info = solve(1.0)                       # un-blended solution of :105-108
        if self.fixed_precision is False:
            # diag(SEyy') = diag(SEyy) - diag(mu invV mu^T) + diag(mu_0 invV_0 mu_0^T): K x n x p' x p' multiply-adds
            d1 = ((mu1 @ invV1) * mu1).sum(-1)
            d0 = ((s["mu0"] @ s["invV0"]) * s["mu0"]).sum(-1)
            dyy = f(SEyy.expand(full + (n, n)), dev).reshape(C, n, n).diagonal(dim1=-2, dim2=-1) - d1 + d0
            self.invU.ss_update(dyy.view(full + (n,)), Nf.view(full).unsqueeze(-1), lr=lr)
            if self.uniform_precision is True:
                self.invU.gamma.alpha = self.invU.gamma.alpha.sum(-1, keepdim=True)      # (the reference's own hack, :125-126)
        if float(lr) != 1.0:
            mu1, invV1, V1, ld1, info = solve(lr)                    # blended + symmetrised invV, its inverse and logdet
        self.mu = mu1.view(full + (n, pp))
        self.invV = invV1.view(full + (pp, pp))
        self.V = V1.view(full + (pp, pp))
        self.logdetinvV = ld1.view(full)
        self.info = info

    def KLqprior(self):
        """transforms/MatrixNormalGamma.py:206-225 (no X_mask)."""
        n, p = self.n, self.p
        KL = n / 2.0 * self.logdetinvV - n / 2.0 * self.logdetinvV_0 - n * p / 2.0
        KL = KL + 0.5 * n * (self.invV_0 * self.V).sum(-1).sum(-1)
        dm = self.mu - self.mu_0
        temp = dm.transpose(-2, -1) @ (self.invU.gamma.mean().unsqueeze(-1) * dm)
        KL = KL + 0.5 * (self.invV_0 * temp).sum(-1).sum(-1)
        for i in range(self.event_dim - 2):
            KL = KL.sum(-1)
        if self.uniform_precision is True:
            KL = KL + self.invU.KLqprior() / n
        else:
            KL = KL + self.invU.KLqprior()
        for i in range(self.event_dim - 2):
            KL = KL.sum(-1)
        return KL

    # expectation-input methods and predict follow the reference's expressions on this node's own getters
    def update(self, pX, pY, p=None, lr=1.0, beta=None):
        """transforms/MatrixNormalGamma.py:143-172 (reference op order on torch; the raw-data path is the accelerated one)."""
        sample_shape = pX.shape[:-self.event_dim - self.batch_dim]
        sd = tuple(range(len(sample_shape)))
        if p is None:
            w = lambda t: t.sum(sd)                                              # noqa: E731
            N = torch.tensor(float(torch.Size(sample_shape).numel()), device=self.mu.device)
            N = N.expand(self.batch_shape + self.event_shape[:-2])
        else:
            N = p.sum(sd)
            pv = p.view(p.shape + self.event_dim * (1,))
            w = lambda t: (t * pv).sum(sd)                                       # noqa: E731
        SExx, SEyy, SEyx = w(pX.EXXT()), w(pY.EXXT()), w(pY.EX() @ pX.EX().transpose(-2, -1))
        if self.pad_X:
            SEx, SEy = w(pX.EX()), w(pY.EX())
            SExx = torch.cat((SExx, SEx), dim=-1)
            SEx = torch.cat((SEx, N.view(N.shape + (1, 1))), dim=-2)
            SExx = torch.cat((SExx, SEx.transpose(-2, -1)), dim=-2)
            SEyx = torch.cat((SEyx, SEy.expand(SEyx.shape[:-1] + (1,))), dim=-1)
        self.ss_update(SExx, SEyx, SEyy, N, lr=lr, beta=beta)

    def Elog_like_given_pX_pY(self, pX, pY):
        """transforms/MatrixNormalGamma.py:234-249."""
        base = self.Elog_like(pX.mean(), pY.mean())
        p_in = self.p - int(self.pad_X)
        corr = (pY.ESigma() * self.EinvSigma()).sum(-1).sum(-1) + (pX.ESigma() * self.EXTinvUX()[..., :p_in, :p_in]).sum(-1).sum(-1)
        for i in range(self.event_dim - 2):
            corr = corr.sum(-1)
        return base - 0.5 * corr

    def _predict_factors(self, logprior=None):
        raise NotImplementedError("the fused predict of MixtureofLinearTransforms is built for type='Wishart'")

    # ---- K-sized expectations that differ from MatrixNormalWishart (transforms/MatrixNormalGamma.py:420-471) -----------
    def EinvUX(self):
        return self.invU.gamma.mean().unsqueeze(-1) * self.mu

    def EXTAX(self, A):
        return self.V * (self.invU.gamma.meaninv() * A.diagonal(dim1=-2, dim2=-1)).sum(-1) + self.mu.transpose(-2, -1) @ A @ self.mu

    def EXmMUTAXmMU(self, A):
        return self.V * (self.invU.gamma.meaninv() * A.diagonal(dim1=-2, dim2=-1)).sum(-1).sum(-1)

    def EXTinvUX(self):
        return self.n * self.V + self.mu.transpose(-1, -2) @ (self.invU.gamma.mean().unsqueeze(-1) * self.mu)

    def EXTX(self):
        return self.V * self.invU.gamma.meaninv().sum() + self.mu.transpose(-1, -2) @ self.mu

    def ElogdetinvU(self):
        return self.invU.gamma.loggeomean().sum(-1)

    def ElogdetinvSigma(self):
        return self.invU.gamma.loggeomean().sum(-1)

    def EinvSigma(self):
        return self.invU.mean()

    def ESigma(self):
        return self.invU.ESigma()
