"""MatrixNormalWishart node with the reference's interface for the VB-EM hot path
(transforms/MatrixNormalWishart.py:8-471): Elog_like / raw_update / ss_update / KLqprior, the expectation-input
update(pX, pY, p) / Elog_like_given_pX_pY, predict and the K-sized expectation getters.  The mask / X_mask branches
and the message-passing methods (forward / backward / Elog_like_X ...) are outside the scope table:
the class pyvbmp_b200.install() binds into a pyVBMP tree inherits them from the reference; stand-alone they are absent.
"""
from __future__ import annotations

import math

import torch

from . import _lib, _shapes
from .wishart import Wishart


class MatrixNormalWishart():

    def __init__(self, event_shape, batch_shape=(), prior_parms={'mu': torch.tensor(0.0)}, scale=1.0, mask=None,
                 X_mask=None, pad_X=False, fixed_precision=False):
        """transforms/MatrixNormalWishart.py:20-70 (mu = randn / sqrt(p') + mu_0 consumes the global RNG)."""
        if mask is not None or X_mask is not None:
            raise NotImplementedError("mask / X_mask are outside the accelerated path (SURVEY.md §2.1 #5)")
        self.n = event_shape[-2]
        self.p = event_shape[-1]
        self.pad_X = pad_X
        self.fixed_precision = fixed_precision
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
        self.mu = torch.randn_like(mu_0, requires_grad=False) / torch.sqrt(torch.tensor(self.p, requires_grad=False)) + mu_0

        mshape = batch_shape + event_shape[:-2]
        self.invV_0 = torch.eye(self.p, requires_grad=False).expand(mshape + (self.p, self.p))
        self.invV = self.invV_0
        self.V = self.invV_0                       # inverse of the identity
        self.logdetinvV = torch.zeros(mshape)
        self.logdetinvV_0 = torch.zeros(mshape)

        self.invU = Wishart(event_shape=event_shape[:-2] + (self.n, self.n), batch_shape=batch_shape, scale=scale)

        self.SEyy = 0.0
        self.SExx = 0.0
        self.SEyx = 0.0
        self.N = 0.0
        self.log2pi = torch.tensor(2 * torch.pi, requires_grad=False).log()

    def to_event(self, n):
        """transforms/MatrixNormalWishart.py:72-80."""
        if n == 0:
            return self
        self.event_dim = self.event_dim + n
        self.batch_dim = self.batch_dim - n
        self.event_shape = self.batch_shape[-n:] + self.event_shape
        self.batch_shape = self.batch_shape[:-n]
        self.invU.to_event(n)
        return self

    def to(self, device):
        for k in ("mu_0", "mu", "invV_0", "invV", "V", "logdetinvV", "logdetinvV_0", "log2pi"):
            setattr(self, k, getattr(self, k).to(device))
        for k in ("SEyy", "SExx", "SEyx", "N"):
            if isinstance(getattr(self, k), torch.Tensor):
                setattr(self, k, getattr(self, k).to(device))
        self.invU.to(device)
        return self

    # ---- helpers ------------------------------------------------------------------------------------
    def _full(self):
        full = tuple(self.batch_shape) + tuple(self.event_shape[:-2])
        return full, int(math.prod(full))

    def _state_flat(self):
        full, C = self._full()
        n, pp, dev = self.n, self.p, self.mu.device
        f = _lib.f32
        w = self.invU
        return dict(
            C=C,
            mu0=f(self.mu_0.expand(full + (n, pp)), dev).reshape(C, n, pp), mu=f(self.mu.expand(full + (n, pp)), dev).reshape(C, n, pp),
            invV0=f(self.invV_0.expand(full + (pp, pp)), dev).reshape(C, pp, pp), invV=f(self.invV.expand(full + (pp, pp)), dev).reshape(C, pp, pp),
            V=f(self.V.expand(full + (pp, pp)), dev).reshape(C, pp, pp),
            ldV=f(self.logdetinvV.expand(full), dev).reshape(C), ldV0=f(self.logdetinvV_0.expand(full), dev).reshape(C),
            invU0=f(w.invU_0.expand(full + (n, n)), dev).reshape(C, n, n), invU=f(w.invU.expand(full + (n, n)), dev).reshape(C, n, n),
            U=f(w.U.expand(full + (n, n)), dev).reshape(C, n, n),
            nu0=f(w.nu_0.expand(full), dev).reshape(C), nu=f(w.nu.expand(full), dev).reshape(C),
            ldU=f(w.logdet_invU.expand(full), dev).reshape(C), ldU0=f(w.logdet_invU_0.expand(full), dev).reshape(C),
        )

    def _plan(self, X):
        """X: sample + batch* + extra + (p, 1)."""
        nb, ne = self.batch_dim, self.event_dim
        sample_shape = tuple(X.shape[:X.ndim - nb - ne])
        bstar = tuple(X.shape[X.ndim - nb - ne:X.ndim - ne])
        return _shapes.make_plan(self.batch_shape, self.event_shape[:-2], bstar, sample_shape)

    def _prep(self, plan, logprior=None):
        nb, nx, dev = self.batch_dim, self.event_dim - 2, self.mu.device
        full, C = self._full()
        n, pp = self.n, self.p
        f = _lib.f32
        w = self.invU
        tk = lambda t, tail: _shapes.theta_to_GK(t, plan, nb, nx, tail)   # noqa: E731
        invU = tk(f(w.invU.expand(full + (n, n)), dev), 2)
        nu = tk(f(w.nu.expand(full), dev), 0)
        mu = tk(f(self.mu.expand(full + (n, pp)), dev), 2)
        invV = tk(f(self.invV.expand(full + (pp, pp)), dev), 2)
        lp = None if logprior is None else tk(f(logprior.expand(full), dev), 0)
        D = n + pp - int(self.pad_X)
        Dp = _lib.pad_dim(D)
        return _lib.mnw_prep(invU, nu, mu, invV, lp, C, n, pp, self.pad_X, Dp) + (Dp,)

    def _cols(self, X, Y, plan):
        dev = self.mu.device
        p_in = self.p - int(self.pad_X)
        Xc = _lib.f32(X, dev).reshape(plan.N, plan.GX, p_in)
        Yc = _lib.f32(Y, dev).reshape(plan.N, plan.GX, self.n)
        return Xc, Yc

    # ---- reference protocol ---------------------------------------------------------------------------
    def ss_update(self, SExx, SEyx, SEyy, N, lr=1.0, beta=None):
        """transforms/MatrixNormalWishart.py:82-141, no-mask branch -> vbmp_mnw_update."""
        assert (SExx.ndim == self.batch_dim + self.event_dim)
        assert (SEyx.ndim == self.batch_dim + self.event_dim)
        assert (SEyy.ndim == self.batch_dim + self.event_dim)
        assert (N.ndim == self.batch_dim + self.event_dim - 2)

        if beta is not None:
            self.SExx = beta * self.SExx + SExx
            self.SEyx = beta * self.SEyx + SEyx
            self.SEyy = beta * self.SEyy + SEyy
            self.N = beta * self.N + N
            SExx = self.SExx
            SEyx = self.SEyx
            SEyy = self.SEyy
            N = self.N
        s = self._state_flat()
        full, C = self._full()
        n, pp, dev = self.n, self.p, self.mu.device
        f = _lib.f32
        mu, invV, V, ldV, invU, nu, U, ldU, info = _lib.mnw_update(
            f(SExx.expand(full + (pp, pp)), dev).reshape(C, pp, pp), f(SEyx.expand(full + (n, pp)), dev).reshape(C, n, pp),
            f(SEyy.expand(full + (n, n)), dev).reshape(C, n, n),
            f(torch.as_tensor(N, dtype=torch.float32, device=dev).expand(full), dev).reshape(C),
            s["mu0"], s["invV0"], s["invU0"], s["nu0"], s["mu"], s["invV"], s["invU"], s["nu"],
            C, n, pp, float(lr), self.fixed_precision is not False)
        self.mu = mu.view(full + (n, pp))
        self.invV = invV.view(full + (pp, pp))
        self.V = V.view(full + (pp, pp))
        self.logdetinvV = ldV.view(full)
        if self.fixed_precision is False:
            self.invU._set(invU, nu, U, ldU, info)
        self.info = info

    def _gram(self, X, Y, p=None):
        """K3: weighted Gram statistics of z = [x; y; 1] in kernel layout (G, K, D+1, D+1)."""
        plan = self._plan(X)
        dev = self.mu.device
        D = self.p - int(self.pad_X) + self.n
        Xc, Yc = self._cols(X, Y, plan)
        pc = None if p is None else _lib.f32(p, dev).reshape(plan.N, plan.GP, plan.K)
        G = _lib.gram(Xc, Yc, plan.N, plan.GX, _shapes.idx_tensor(plan.xg, dev), pc, plan.GP,
                      _shapes.idx_tensor(plan.pg, dev), plan.G, plan.K, _lib.pad_dim(D))
        return G, plan

    def _update_from_gram(self, G, plan, weighted, lr=1.0, beta=None, n_samples=None):
        """Blocks of the Gram matrix are SExx / SEyx / SEyy (+ SEx, SEy, N when pad_X) of
        transforms/MatrixNormalWishart.py:185-202."""
        dev = self.mu.device
        n, pp = self.n, self.p
        p_in = pp - int(self.pad_X)
        D = p_in + n
        G = _shapes.GK_to_theta(G, plan, (D + 1, D + 1))
        SEyy = G[..., p_in:D, p_in:D]
        if self.pad_X:
            ix = list(range(p_in)) + [D]
            SExx = G[..., ix, :][..., :, ix]
            SEyx = G[..., p_in:D, :][..., :, ix]
        else:
            SExx = G[..., :p_in, :p_in]
            SEyx = G[..., p_in:D, :p_in]
        if not weighted:
            nn = float(plan.N if n_samples is None else n_samples)
            N = torch.tensor(nn, device=dev).expand(self.batch_shape + self.event_shape[:-2])
        else:
            N = G[..., D, D]
        self.ss_update(SExx, SEyx, SEyy, N, lr=lr, beta=beta)

    def raw_update(self, X, Y, p=None, lr=1.0, beta=None):
        """transforms/MatrixNormalWishart.py:174-204: one weighted Gram pass over z = [x; y; 1]."""
        G, plan = self._gram(X, Y, p)
        self._update_from_gram(G, plan, p is not None, lr, beta)

    def _beliefs_plan(self, pX, pY):
        """Beliefs whose means look like raw data to the kernels: (sample..., 1, p, 1) / (sample..., 1, n, 1) on the device,
        one component axis.  Returns (N, means, flattened covariances) or None."""
        mx, my = pX.EX(), pY.EX()
        p_in = self.p - int(self.pad_X)
        if not (self.batch_dim == 1 and self.event_dim == 2 and mx.is_cuda and mx.ndim >= 4 and mx.shape[-3] == 1
                and my.shape[-3] == 1 and mx.shape[:-3] == my.shape[:-3]):
            return None
        sample = tuple(mx.shape[:-3])
        N = 1
        for v in sample:
            N *= v
        # a point mass (dists/Delta.py: no ESigma) contributes no covariance term; a covariance shared by all samples
        # stays one row (its weighted sum is NA_k Sigma, its trace term one K-vector) instead of N copies
        def flat(b, d):
            if not hasattr(b, "ESigma"):
                return None
            S = b.ESigma()
            if S.numel() == d * d:
                return S.reshape(1, d * d)
            return S.expand(sample + (1, d, d)).reshape(N, d * d)
        return N, sample, mx, my, flat(pX, p_in), flat(pY, self.n)

    def update(self, pX, pY, p=None, lr=1.0, beta=None):
        """transforms/MatrixNormalWishart.py:143-172: M-step from Gaussian beliefs about inputs and outputs.
        E[xx^T] = Sigma_x + mu_x mu_x^T (same for y; the cross term uses the means only), so the statistics are the weighted
        Gram pass (K3) over the MEANS plus the responsibility-weighted sums of the covariances — one skinny
        (K x N) (N x p^2) product per belief, added into the x-x and y-y blocks before the update kernel."""
        bp = self._beliefs_plan(pX, pY) if p is not None else None
        if bp is not None:
            N, sample, mx, my, Sx, Sy = bp
            G, plan = self._gram(mx, my, p)
            if plan.G == 1 and plan.GX == 1:
                K, n = plan.K, self.n
                p_in = self.p - int(self.pad_X)
                D = p_in + n
                P2 = _lib.f32(p, self.mu.device).reshape(N, K)
                Gk = G.view(K, D + 1, D + 1)
                # sum_n p[n,k] Sigma_n: (K x N) (N x p^2) over the sample axis (vbmp_wsum, the Gram kernel's "lin" mode);
                # shapes outside its window (tiny N, K % 4 != 0) are K-sized-output matmuls on the device
                def wsum(Sf):
                    Sf = _lib.f32(Sf)
                    if Sf.shape[0] == 1 and N > 1:
                        return P2.sum(0)[:, None] * Sf
                    if _lib.wsum_supported(N, K, Sf.shape[1], Sf.stride(0)) and Sf.data_ptr() % 16 == 0:
                        return _lib.wsum(P2, Sf)
                    return P2.t() @ Sf
                if Sx is not None:
                    Gk[:, :p_in, :p_in] += wsum(Sx).view(K, p_in, p_in)
                if Sy is not None:
                    Gk[:, p_in:D, p_in:D] += wsum(Sy).view(K, n, n)
                self._update_from_gram(G, plan, p is not None, lr, beta)
                return
        # any other layout: the reference's op order on torch
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
        """transforms/MatrixNormalWishart.py:234-249.  The expected log likelihood is linear in E[xx^T], E[yy^T], so it is
        the ordinary Elog_like at the means (K1 + K2) minus 1/2 tr(Sigma_y E[invSigma_k]) + 1/2 tr(Sigma_x E[X^T invU X]_k)
        (the x-x block), two skinny products over the flattened covariances."""
        base = self.Elog_like(pX.mean(), pY.mean())
        p_in = self.p - int(self.pad_X)
        Exx = self.EXTinvUX()[..., :p_in, :p_in]
        bp = self._beliefs_plan(pX, pY)
        if bp is not None and self.event_dim == 2:
            N, sample, mx, my, Sx, Sy = bp
            K = self.batch_shape[0]
            # tr(Sigma_y E[invSigma_k]) + tr(Sigma_x E[X^T invU X]_k): two (N x n^2) (n^2 x K) products (vbmp_rowterm;
            # vbmp_rowgemm for the shapes outside its window)
            corr = {}                          # per-sample rows (N, K) and / or one row shared by all samples (1, K)
            base2d = base.view(N, K) if base.is_contiguous() else None
            def add(Sf, Lm):
                nonlocal base2d
                Sf, Lm = _lib.f32(Sf), _lib.f32(Lm.reshape(K, -1).t().contiguous())
                key = "shared" if (Sf.shape[0] == 1 and N > 1) else "rows"
                if key == "rows" and base2d is not None and _lib.rowterm_supported(N, Sf.shape[1], K, Sf.stride(0)):
                    _lib.rowterm(Sf, Lm, C=base2d, alpha=-0.5, accumulate=True)      # base -= 1/2 Sf Lm, in place (tcgen05)
                    return
                corr[key] = _lib.rowgemm(Sf, Lm, out=corr.get(key), accumulate=key in corr)
            if Sy is not None:
                add(Sy, self.EinvSigma().expand(K, self.n, self.n))
            if Sx is not None:
                add(Sx, Exx.expand(K, p_in, p_in))
            if "rows" in corr:
                base = base - 0.5 * corr["rows"].view(sample + (K,))
            if "shared" in corr:
                base = base - 0.5 * corr["shared"].view(len(sample) * (1,) + (K,))
            return base
        corr = 0.0
        if hasattr(pY, "ESigma"):
            corr = corr + (pY.ESigma() * self.EinvSigma()).sum(-1).sum(-1)
        if hasattr(pX, "ESigma"):
            corr = corr + (pX.ESigma() * Exx).sum(-1).sum(-1)
        if isinstance(corr, float):
            return base
        for i in range(self.event_dim - 2):
            corr = corr.sum(-1)
        return base - 0.5 * corr

    def KLqprior(self):
        """transforms/MatrixNormalWishart.py:206-216 -> vbmp_mnw_kl."""
        s = self._state_flat()
        full, C = self._full()
        KL = _lib.mnw_kl(s["mu0"], s["mu"], s["invV0"], s["V"], s["ldV"], s["ldV0"], s["invU0"], s["U"], s["nu0"],
                         s["nu"], s["ldU"], s["ldU0"], C, self.n, self.p).view(full)
        for i in range(self.event_dim - 2):
            KL = KL.sum(-1)
        return KL

    def Elog_like(self, X, Y):
        """transforms/MatrixNormalWishart.py:219-232 -> K1 + K2 with z = [x; y]."""
        plan = self._plan(X)
        dev = self.mu.device
        W, m, cst, info, Dp = self._prep(plan)
        Xc, Yc = self._cols(X, Y, plan)
        out = _lib.estep(Xc, Yc, plan.N, plan.GX, _shapes.idx_tensor(plan.xg, dev), W, m, cst, plan.G, plan.K, Dp, 0)
        out = _shapes.logits_to_ref(out, plan)
        for i in range(self.event_dim - 2):
            out = out.sum(-1)
        return out

    # The message-passing methods (transforms/MatrixNormalWishart.py:251-398: Elog_like_X, forward, backward, ...) are not
    # defined here on purpose: the class install() binds into a pyVBMP tree inherits them from the reference class
    # (pyvbmp_b200/install.py), and this stand-alone class says where they live instead of running something else.
    _REFERENCE_ONLY = ("Elog_like_X", "Elog_like_X_given_pY", "Elog_like_X_given_Y", "Eforward", "forward", "backward",
                       "postdict", "predict_given_pX", "Ebackward", "forward_old")

    def __getattr__(self, name):
        if name in MatrixNormalWishart._REFERENCE_ONLY:
            raise AttributeError(f"MatrixNormalWishart.{name} is a message-passing method outside the VB-EM hot path "
                                 "(SURVEY.md §2.1 #5); pyvbmp_b200.install() keeps the reference's implementation of it")
        raise AttributeError(f"{type(self).__name__!r} object has no attribute {name!r}")

    # ---- K-sized expectations (transforms/MatrixNormalWishart.py:400-471) ------------------------------
    def mean(self):
        return self.mu

    def bias(self):
        return self.mu[..., -1:] if self.pad_X is True else torch.tensor(0.0)

    def weights(self):
        return self.mu[..., :-1] if self.pad_X is True else self.mu

    def var(self):
        return self.ESigma().diagonal(dim1=-1, dim2=-2).unsqueeze(-1) * self.V.diagonal(dim1=-1, dim2=-2).unsqueeze(-2)

    def predict(self, X):
        """transforms/MatrixNormalWishart.py:381-390: natural parameters of p(y | x) per component and the per-component
        log evidence.  Same op order as the reference (small per-component matrices, one batched product with X)."""
        from .mvn import MultivariateNormal_vector_format
        if self.pad_X:
            invSigmamu_y = (self.EinvUX()[..., :, :-1] @ X + self.EinvUX()[..., :, -1:])
            Res = (-0.5 * X.transpose(-1, -2) @ self.EXTinvUX()[..., :-1, :-1] @ X - self.EXTinvUX()[..., -1:, :-1] @ X
                   - 0.5 * self.EXTinvUX()[..., -1:, -1:])
        else:
            invSigmamu_y = (self.EinvUX() @ X)
            Res = -0.5 * X.transpose(-1, -2) @ self.EXTinvUX() @ X
        Res = Res.squeeze(-1).squeeze(-1) + 0.5 * self.ElogdetinvSigma() - 0.5 * self.n * self.log2pi
        pY = MultivariateNormal_vector_format(invSigma=self.EinvSigma(), invSigmamu=invSigmamu_y)
        return pY, Res - pY.Res()

    def _predict_factors(self, logprior=None):
        """Whitened form of `predict`'s per-component log evidence for the E-step kernel (K2):
        Res - pY.Res() = -1/2 n xt^T V xt + 1/2 (E logdet invSigma - logdet E invSigma),  xt = [x;1]  (the mu-dependent
        quadratic terms of :384 and of MultivariateNormal_vector_format.Res cancel), so with n V = U U^T, U upper
        triangular, it is cst - 1/2 ||W^T x - m||^2 with W = U[:p], m = -U[p] (the row of the constant feature).
        K small matrices, fp64 on the device; returns (W (K,Dp,Dp), m (K,Dp), cst (K), Dp)."""
        assert self.batch_dim == 1 and self.event_dim == 2
        dev = self.mu.device
        pp, K = self.p, self.batch_shape[0]
        p_in = pp - int(self.pad_X)
        A = (self.n * self.V).double().expand(K, pp, pp)
        U = torch.linalg.cholesky(A.flip(-1, -2)).flip(-1, -2)            # A = U U^T, U upper triangular
        Dp = _lib.pad_dim(pp)
        W = torch.zeros(K, Dp, Dp, dtype=torch.float64, device=dev)
        m = torch.zeros(K, Dp, dtype=torch.float64, device=dev)
        W[:, :p_in, :pp] = U[:, :p_in, :]
        if self.pad_X:
            m[:, :pp] = -U[:, p_in, :]
        cst = 0.5 * (self.ElogdetinvSigma().double() - self.EinvSigma().double().logdet()).expand(K)
        if logprior is not None:
            cst = cst + logprior.double()
        return W.float().contiguous(), m.float().contiguous(), cst.float().contiguous(), Dp

    def EinvUX(self):
        return self.invU.EinvSigma() @ self.mu

    def EXTinvU(self):
        return self.mu.transpose(-2, -1) @ self.invU.EinvSigma()

    def EXTAX(self, A):
        return self.V * (self.invU.ESigma() * A).sum(-1).sum(-1) + self.mu.transpose(-2, -1) @ A @ self.mu

    def EXmMUTAXmMU(self, A):
        return self.V * (self.invU.ESigma() * A).sum(-1).sum(-1)

    def EXAXT(self, A):
        return self.ESigma() * (self.V * A).sum(-1).sum(-1) + self.mu @ A @ self.mu.transpose(-2, -1)

    def EXmMUAXmMUT(self, A):
        return self.ESigma() * (self.V * A).sum(-1).sum(-1)

    def EXTinvUX(self):
        return self.n * self.V + self.mu.transpose(-1, -2) @ self.invU.EinvSigma() @ self.mu

    def EXinvVXT(self):
        return self.p * self.invU.ESigma() + self.mu @ self.invV @ self.mu.transpose(-1, -2)

    def EXmMUTinvUXmMU(self):
        return self.n * self.V

    def EXmMUinvVXmMUT(self):
        return self.p * self.invU.ESigma()

    def EXTX(self):
        return self.V * self.invU.ESigma().diagonal().sum() + self.mu.transpose(-1, -2) @ self.mu

    def EXXT(self):
        return self.V.diagonal().sum() * self.invU.ESigma() + self.mu @ self.mu.transpose(-1, -2)

    def ElogdetinvU(self):
        return self.invU.ElogdetinvSigma()

    def logdetEinvSigma(self):
        return self.invU.logdetEinvSigma()

    def ElogdetinvSigma(self):
        return self.invU.ElogdetinvSigma()

    def EinvSigma(self):
        return self.invU.EinvSigma()

    def invEinvSigma(self):
        return self.invU.invEinvSigma()

    def ESigma(self):
        return self.invU.ESigma()
