"""MixtureofLinearTransforms with the reference's interface for the raw-data VB-EM loop
(transforms/MixtureofLinearTransforms.py:10-213, type='Wishart').  update_assignments is the fused
K1 + K2 (softmax epilogue) sequence; the M-step is one weighted Gram pass + the MNW update kernel."""
from __future__ import annotations

import torch

from . import _lib, _shapes, sharding
from .dirichlet import Dirichlet
from .mnw import MatrixNormalWishart
from .mng import MatrixNormalGamma


def fused_update_assignments(self, X, Y, fallback=None):
    """transforms/MixtureofLinearTransforms.py:34-41 on the CUDA path (also the install() method patch; ``fallback`` = the
    reference's own method, taken for anything this does not fuse)."""
    W = self.W
    other = fallback if fallback is not None else generic_update_assignments
    if (not isinstance(W, MatrixNormalWishart) or self.batch_dim != 0 or W.event_dim != 2
            or (fallback is not None and not W.mu.is_cuda)):
        return other(self, X, Y)
    Xv, Yv = X.unsqueeze(-3), Y.unsqueeze(-3)
    plan = W._plan(Xv)
    dev = W.mu.device
    Wt, m, cst, info, Dp = W._prep(plan, logprior=self.pi.loggeomean())
    Xc, Yc = W._cols(Xv, Yv, plan)
    p, logZn, NA, logZ = _lib.estep(Xc, Yc, plan.N, plan.GX, _shapes.idx_tensor(plan.xg, dev), Wt, m, cst,
                                    plan.G, plan.K, Dp, 1)
    self.p = p.view(plan.sample_shape + (plan.K,))
    self.logZ = logZn.view(plan.sample_shape)          # per-sample log normaliser (:41)
    self.NA = NA.view(plan.K)                          # = p.sum(0) for a single sample dim


def generic_update_assignments(self, X, Y):
    log_p = self.W.Elog_like(X.unsqueeze(-3), Y.unsqueeze(-3)) + self.pi.loggeomean()
    self.logZ = torch.logsumexp(log_p, -1)
    self.p = (log_p - self.logZ.unsqueeze(-1)).exp()
    self.NA = None


class MixtureofLinearTransforms():

    def __init__(self, n, p, dim, batch_shape=(), pad_X=True, type='Wishart'):
        """transforms/MixtureofLinearTransforms.py:12-32."""
        self.n = n
        self.p = p
        self.dim = dim
        self.event_dim = 1
        self.event_shape = (dim,)
        self.batch_dim = len(batch_shape)
        self.batch_shape = batch_shape
        if type == 'Wishart':
            self.W = MatrixNormalWishart(event_shape=(n, p), batch_shape=batch_shape + (dim,),
                                         scale=1.0 / dim ** (1.0 / n), pad_X=pad_X)
        elif type == 'Gamma':
            self.W = MatrixNormalGamma(event_shape=(n, p), batch_shape=batch_shape + (dim,),
                                       scale=1.0 / dim ** (1.0 / n), pad_X=pad_X)
        else:
            raise ValueError('type must be either Wishart (default) or Gamma')
        self.pi = Dirichlet(event_shape=(dim,), batch_shape=batch_shape)
        self.KL_last = None
        self.ELBO_last = -torch.tensor(torch.inf)

    def to(self, device):
        self.W.to(device)
        self.pi.to(device)
        self.ELBO_last = self.ELBO_last.to(device)
        return self

    update_assignments = fused_update_assignments

    def raw_update(self, X, Y, iters=1, lr=1.0, verbose=False):
        """transforms/MixtureofLinearTransforms.py:50-61."""
        for i in range(iters):
            self.update_assignments(X, Y)
            NA = self.NA if (getattr(self, "NA", None) is not None and self.p.ndim == 2) else self.p.sum(0)
            if sharding.enabled():
                G, plan = self.W._gram(X.unsqueeze(-3), Y.unsqueeze(-3), self.p)
                G, logZ_sum, NA = sharding.all_reduce_packed([G, self.logZ.sum(0), NA])
                ELBO = logZ_sum - self.KLqprior()
                self.pi.ss_update(NA, lr=lr)
                self.W._update_from_gram(G, plan, True, lr)
            else:
                ELBO = self.ELBO()
                self.pi.ss_update(NA, lr=lr)
                self.W.raw_update(X.unsqueeze(-3), Y.unsqueeze(-3), p=self.p, lr=lr)
            if verbose:
                print('MixLinearTransform: Percent Change in ELBO = ', ((ELBO - self.ELBO_last) / self.ELBO_last.abs()).data * 100)
            self.ELBO_last = ELBO

    def update_assignments_given_pX_pY(self, pX, pY):
        """transforms/MixtureofLinearTransforms.py:62-69: max-shifted softmax of the expected log likelihoods."""
        ELL = self.W.Elog_like_given_pX_pY(pX.unsqueeze(-3), pY.unsqueeze(-3))
        if ELL.is_cuda and ELL.ndim == 2 and ELL.is_contiguous() and ELL.dtype == torch.float32 and self.batch_dim == 0:
            # one pass: + loggeomean, logsumexp, responsibilities (in place over the logits), NA (vbmp_softmax_rows)
            self.p, self.logZ, self.NA, _ = _lib.softmax_rows(ELL, colbias=_lib.f32(self.pi.loggeomean()), out=ELL)
            return
        log_p = ELL + self.pi.loggeomean()
        self.logZ = torch.logsumexp(log_p, -1)
        self.p = (log_p - self.logZ.unsqueeze(-1)).exp()
        self.NA = None

    def Elog_like_given_pX_pY(self, pX, pY):
        """transforms/MixtureofLinearTransforms.py:71-75."""
        ELL = (self.W.Elog_like(pX.unsqueeze(-3), pY.unsqueeze(-3)) * self.p).sum(-1)
        for i in range(self.event_dim - 1):
            ELL = ELL.sum(-1)
        return ELL

    def update(self, pX, pY, iters=1, lr=1, verbose=False):
        """transforms/MixtureofLinearTransforms.py:77-90: VB-EM on Gaussian beliefs about inputs and outputs
        (SURVEY.md §8f #2)."""
        if sharding.enabled():
            raise NotImplementedError("update(pX, pY) is not sample-sharded; use raw_update for sharded runs")
        for i in range(iters):
            self.update_assignments_given_pX_pY(pX, pY)
            ELBO = self.ELBO()
            self.pi.ss_update(self.NA if (self.NA is not None and self.p.ndim == 2) else self.p.sum(0), lr=lr)
            self.W.update(pX.unsqueeze(-3), pY.unsqueeze(-3), p=self.p, lr=lr)
            if verbose:
                print('MixLinearTransform: Percent Change in ELBO = ', ((ELBO - self.ELBO_last) / self.ELBO_last.abs()).data * 100)
            self.ELBO_last = ELBO

    PREDICT_ROWS = 1 << 17       # rows per block of the (rows, K, n) temporaries of the moment sums (1 GiB at K n = 2048)

    def predict(self, X):
        """transforms/MixtureofLinearTransforms.py:91-108: mixture-of-experts predictive distribution of Y and the gate
        probabilities for inputs X (..., p, 1).  On the CUDA path the gate probabilities come from the fused E-step kernel
        (K2, softmax epilogue) on the whitened form of the per-component evidence (MatrixNormalWishart._predict_factors);
        the component means and sum_k p_k ESigma_k are one product each per block of rows (shared operands: vbmp_rowgemm_ex,
        tcgen05), the per-sample weighted rank-K update is vbmp_moe_moments (one warp per sample)."""
        from .mvn import MultivariateNormal_vector_format
        W = self.W
        if not (isinstance(W, MatrixNormalWishart) and not isinstance(W, MatrixNormalGamma) and self.batch_dim == 0
                and W.event_dim == 2 and X.is_cuda and X.ndim == 3):
            pY, Res = W.predict(X.unsqueeze(-3))
            log_p = Res + self.pi.loggeomean()
            log_p = log_p - log_p.max(-1, True)[0]
            p = log_p.exp()
            p = p / p.sum(-1, True)
            pe = p.unsqueeze(-1).unsqueeze(-1)
            Sigma = ((pY.ESigma() + pY.mean() @ pY.mean().transpose(-2, -1)) * pe).sum(-3)
            mu = (pY.mean() * pe).sum(-3)
            Sigma = Sigma - mu @ mu.transpose(-2, -1)
            return MultivariateNormal_vector_format(mu=mu, Sigma=Sigma), p
        dev = W.mu.device
        N, K, n = X.shape[0], self.dim, self.n
        p_in = W.p - int(W.pad_X)
        Wt, m, cst, Dp = W._predict_factors(self.pi.loggeomean())
        Xc = _lib.f32(X, dev).reshape(N, 1, p_in)
        xg = _shapes.idx_tensor((0,), dev)
        p = _lib.estep(Xc, None, N, 1, xg, Wt, m, cst, 1, K, Dp, 1)[0].view(N, K)
        M = W.mu                                                           # (K, n, p'): mean_k(x) = M_k [x;1]
        ES = W.EinvSigma().inverse().expand(K, n, n)                       # = pY.ESigma(): (E invSigma)^-1 = invU / nu (:389)
        Mw = M[..., :p_in].reshape(K * n, p_in).t().contiguous()           # (p, K n)
        Mb = M[..., -1].reshape(K * n).contiguous() if W.pad_X else None
        ESf = ES.reshape(K, n * n).contiguous()
        mu = torch.empty(N, n, 1, device=dev)
        Sigma = torch.empty(N, n, n, device=dev)
        X2 = Xc.view(N, p_in)
        # per block of rows (fp32): all component means at once ((rows x p') (p' x K n), the bias folded in as the product's
        # additive term) and sum_k p_k ESigma_k ((rows x K) (K x n^2)) are one GEMM each with a shared operand; the per-sample
        # weighted rank-K update is vbmp_moe_moments, which writes mu and Sigma in place (Sigma starts as the base term)
        for a in range(0, N, self.PREDICT_ROWS):
            b = min(N, a + self.PREDICT_ROWS)
            pe = p[a:b]
            S = Sigma[a:b]
            mean = _lib.rowgemm(X2[a:b], Mw, bias=Mb)                                           # (rows, K n)
            _lib.rowgemm(pe, ESf, out=S.view(b - a, n * n))                                     # sum_k p_k ESigma_k
            if n <= 32:
                _lib.moe_moments(mean, pe, S, b - a, K, n, mu=mu[a:b].view(b - a, n), Sigma=S)
                continue
            mean = mean.view(b - a, K, n)                                                       # n > 32: batched torch
            mu_b = torch.bmm(pe.unsqueeze(1), mean).squeeze(1)                                  # (rows, n)
            A = mean * pe.sqrt().unsqueeze(-1)
            S.baddbmm_(A.transpose(1, 2), A)
            mu[a:b, :, 0] = mu_b
            S.sub_(mu_b.unsqueeze(-1) * mu_b.unsqueeze(-2))
        return MultivariateNormal_vector_format(mu=mu, Sigma=Sigma), p

    def KLqprior(self):
        return self.pi.KLqprior() + self.W.KLqprior().sum(-1)

    def ELBO(self):
        """transforms/MixtureofLinearTransforms.py:126-130."""
        logZ = self.logZ.sum(0)
        while logZ.ndim > self.batch_dim:
            logZ = logZ.sum(0)
        return logZ - self.KLqprior()

    def assignment_pr(self):
        return self.p

    def assignment(self):
        return self.p.argmax(-1)

    def mean(self):
        return self.p

    def event_average(self, A):
        p = self.p
        for i in range(self.W.event_dim):
            p = p.unsqueeze(-1)
        out = (A * p)
        for i in range(self.event_dim):
            out = out.sum(-self.W.event_dim - 1)
        return out

    def average(self, A):
        out = self.p * A
        for i in range(self.event_dim):
            out = out.sum(-1)
        return out

    def EinvUX(self):
        return self.event_average(self.W.EinvUX())

    def EXTinvU(self):
        return self.event_average(self.W.EXTinvU())

    def EXTinvUX(self):
        return self.event_average(self.W.EXTinvUX())

    def EXTAX(self, A):
        return self.event_average(self.W.EXTAX(A))

    def EXAXT(self, A):
        return self.event_average(self.W.EXAXT(A))

    def EXinvVXT(self):
        return self.event_average(self.W.EXinvVXT())

    def EXmMUTinvUXmMU(self):
        return self.event_average(self.W.EXmMUTinvUXmMU())

    def EXmMUinvVXmMUT(self):
        return self.event_average(self.W.EXmMUinvVXmMUT())

    def EXTX(self):
        return self.event_average(self.W.EXTX())

    def EXXT(self):
        return self.event_average(self.W.EXXT())

    def ElogdetinvU(self):
        return self.average(self.W.invU.ElogdetinvSigma())

    def EinvSigma(self):
        return self.event_average(self.W.EinvSigma())

    def ESigma(self):
        return self.event_average(self.W.ESigma())

    def ElogdetinvSigma(self):
        return self.average(self.W.ElogdetinvSigma())

    def weights(self):
        return self.W.weights()

    def bias(self):
        return self.W.bias()
