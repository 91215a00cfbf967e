"""CPU oracle for the conjugate VB-EM hot path (TEST INFRASTRUCTURE ONLY).

This module is a functional restatement, in plain CPU torch, of the algorithm the
reference (bayesianempirimancer/pyVBMP) runs on the NIW / MNW E-step + M-step
path.  It is the checker for the CUDA path: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` leg may import it.  Nothing under ``pyvbmp_b200/`` imports it and the
product path never routes through it.

Pinning: the reference has no golden vectors or asserts of its own (SURVEY.md
§8c), so the oracle is pinned against outputs of the reference itself, produced
in the build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference``) and committed as ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` replays every fixture through this module.

State is carried in plain dicts of tensors (no classes) so that nothing here can
be mistaken for the product classes.  Every function cites the reference
file:line whose arithmetic it restates; the ``*_exact`` functions keep the
reference's operation order (broadcast multiply + sum) so fp32 results agree to
rounding, the ``*_fast`` functions restate the same maths with matmuls in the
caller's dtype (fp64 for ground truth) so shapes like N=65 536, d=64, K=256
finish in seconds.
"""
from __future__ import annotations

import math
import torch

# --------------------------------------------------------------------------------------
# Wishart                                   reference: dists/Wishart.py
# --------------------------------------------------------------------------------------

def wishart_new(dim, batch_shape=(), scale=1.0, extra_event=(), dtype=torch.float32):
    """Prior/posterior state of a Wishart node.  dists/Wishart.py:9-26.

    invU_0 = scale^2 I (expanded), nu_0 = dim + 2, U = inv(invU), logdet via LU.
    """
    ev = tuple(extra_event) + (dim, dim)
    invU0 = (scale ** 2 * torch.eye(dim, dtype=dtype)).expand(tuple(batch_shape) + ev)
    nu0 = torch.tensor(dim + 2.0, dtype=dtype).expand(tuple(batch_shape) + tuple(extra_event))
    return {
        "dim": dim,
        "invU_0": invU0, "nu_0": nu0, "logdet_invU_0": invU0.logdet(),
        "invU": invU0, "U": invU0.inverse(), "nu": nu0, "logdet_invU": invU0.logdet(),
        "SExx": 0.0, "N": 0.0,
    }


def _mv_sum(fn, a, dim):
    """sum_{i<dim} fn(a - i/2).  dists/Wishart.py:37-41 (log_mvgamma / log_mvdigamma)."""
    return fn(a.unsqueeze(-1) - torch.arange(dim, dtype=a.dtype) / 2.0).sum(-1)


def wishart_ss_update(w, SExx, N, lr=1.0, beta=None):
    """dists/Wishart.py:43-56: lr-blend of (invU_0+SExx), (nu_0+N); U = inverse; logdet."""
    if beta is not None:
        w["SExx"] = SExx + beta * w["SExx"]
        w["N"] = N + beta * w["N"]
        SExx, N = w["SExx"], w["N"]
    w["invU"] = lr * (w["invU_0"] + SExx) + (1.0 - lr) * w["invU"]
    w["nu"] = lr * (w["nu_0"] + N) + (1.0 - lr) * w["nu"]
    w["U"] = w["invU"].inverse()
    w["logdet_invU"] = w["invU"].logdet()


def wishart_EinvSigma(w):
    """dists/Wishart.py:76-77."""
    return w["U"] * w["nu"].view(w["nu"].shape + (1, 1))


def wishart_ESigma(w):
    """dists/Wishart.py:70-74."""
    return w["invU"] / (w["nu"].view(w["nu"].shape + (1, 1)) - w["dim"] - 1)


def wishart_ElogdetinvSigma(w):
    """dists/Wishart.py:82-83: d log 2 - logdet(invU) + psi_d(nu/2)."""
    d = w["dim"]
    return d * math.log(2.0) - w["logdet_invU"] + _mv_sum(torch.digamma, w["nu"] / 2.0, d)


def wishart_kl(w, extra_event_dims=0):
    """dists/Wishart.py:88-94."""
    d = w["dim"]
    out = w["nu_0"] / 2.0 * (w["logdet_invU"] - w["logdet_invU_0"]) \
        + w["nu"] / 2.0 * (w["invU_0"] * w["U"]).sum(-1).sum(-1) - w["nu"] * d / 2.0
    out = out + _mv_sum(torch.lgamma, w["nu_0"] / 2.0, d) - _mv_sum(torch.lgamma, w["nu"] / 2.0, d) \
        + (w["nu"] - w["nu_0"]) / 2.0 * _mv_sum(torch.digamma, w["nu"] / 2.0, d)
    for _ in range(extra_event_dims):
        out = out.sum(-1)
    return out


# --------------------------------------------------------------------------------------
# Dirichlet                                 reference: dists/Dirichlet.py
# --------------------------------------------------------------------------------------

def dirichlet_new(event_shape, batch_shape=(), alpha0=0.5, dtype=torch.float32):
    """dists/Dirichlet.py:4-11: alpha = alpha_0 (1 + rand)  (consumes the global RNG)."""
    a0 = torch.as_tensor(alpha0, dtype=dtype).expand(tuple(batch_shape) + tuple(event_shape))
    return {"event_dim": len(event_shape), "batch_dim": len(batch_shape),
            "alpha_0": a0, "alpha": a0 * (1.0 + torch.rand(a0.shape, dtype=dtype)), "NA": 0.0}


def dirichlet_ss_update(dr, NA, lr=1.0, beta=None):
    """dists/Dirichlet.py:22-28."""
    dr["NA"] = beta * dr["NA"] + NA if beta is not None else NA
    dr["alpha"] = lr * (dr["NA"] + dr["alpha_0"]) + (1 - lr) * dr["alpha"]


def dirichlet_loggeomean(dr):
    """dists/Dirichlet.py:52-53: psi(alpha) - psi(sum alpha)."""
    ed = list(range(-dr["event_dim"], 0))
    return dr["alpha"].digamma() - dr["alpha"].sum(ed, keepdim=True).digamma()


def dirichlet_kl(dr):
    """dists/Dirichlet.py:63-83 (inf lgamma / -inf digamma entries are zeroed)."""
    ed = list(range(-dr["event_dim"], 0))
    a, a0 = dr["alpha"], dr["alpha_0"]

    def lg(x):
        o = x.lgamma().clone()
        o[o == torch.inf] = 0
        return o

    def dg(x):
        o = x.digamma().clone()
        o[o == -torch.inf] = 0
        return o

    asum, a0sum = a.sum(ed), a0.sum(ed)
    KL = asum.lgamma() - lg(a).sum(ed) - a0sum.lgamma() + lg(a0).sum(ed)
    KL = KL + ((a - a0) * (dg(a) - asum.digamma().view(asum.shape + (1,) * dr["event_dim"]))).sum(ed)
    while KL.ndim > dr["batch_dim"]:
        KL = KL.sum(-1)
    return KL


# --------------------------------------------------------------------------------------
# NormalInverseWishart                      reference: dists/NormalInverseWishart.py
# --------------------------------------------------------------------------------------

def niw_new(event_shape, batch_shape=(), scale=1.0, fixed_precision=False, dtype=torch.float32):
    """dists/NormalInverseWishart.py:6-37: lambda_0=1, mu_0=0, mu = mu_0 + randn (RNG)."""
    event_shape, batch_shape = tuple(event_shape), tuple(batch_shape)
    d = event_shape[-1]
    lam0 = torch.tensor(1.0, dtype=dtype).expand(batch_shape + (len(event_shape) - 1) * (1,))
    mu0 = torch.tensor(0.0, dtype=dtype).expand(batch_shape + event_shape)
    mu = mu0 + torch.randn_like(mu0)
    return {
        "dim": d, "event_shape": event_shape, "batch_shape": batch_shape,
        "event_dim": len(event_shape), "batch_dim": len(batch_shape),
        "fixed_precision": fixed_precision,
        "lambda_mu_0": lam0, "lambda_mu": lam0, "mu_0": mu0, "mu": mu,
        "invU": wishart_new(d, batch_shape, scale, extra_event=event_shape[:-1], dtype=dtype),
        "SExx": torch.tensor(0.0, dtype=dtype), "SEx": torch.tensor(0.0, dtype=dtype),
        "N": torch.tensor(0.0, dtype=dtype),
    }


def niw_EinvSigmamu(s):
    """dists/NormalInverseWishart.py:122-123."""
    return (wishart_EinvSigma(s["invU"]) * s["mu"].unsqueeze(-2)).sum(-1)


def niw_EXTinvUX(s):
    """dists/NormalInverseWishart.py:131-132: nu mu^T U mu + d/lambda."""
    return (s["mu"].unsqueeze(-1) * wishart_EinvSigma(s["invU"]) * s["mu"].unsqueeze(-2)).sum(-1).sum(-1) \
        + s["dim"] / s["lambda_mu"]


def niw_elog_like_exact(s, X):
    """dists/NormalInverseWishart.py:91-97, same op order (materialises (...,K,d,d))."""
    L = wishart_EinvSigma(s["invU"])
    out = -0.5 * ((X.unsqueeze(-1) * L).sum(-2) * X).sum(-1) + (X * niw_EinvSigmamu(s)).sum(-1) \
        - 0.5 * niw_EXTinvUX(s)
    out = out + 0.5 * wishart_ElogdetinvSigma(s["invU"]) - 0.5 * s["dim"] * math.log(2 * math.pi)
    for _ in range(s["event_dim"] - 1):
        out = out.sum(-1)
    return out


def niw_elog_like_fast(s, X2d):
    """Same quantity as niw_elog_like_exact for event_dim == 1, batch (K,), X2d = (N,d),
    restated with matmuls (O(N K d) memory).  Run it in fp64 for ground truth."""
    L = wishart_EinvSigma(s["invU"])                       # (K,d,d)
    XL = torch.einsum("ni,kij->nkj", X2d, L)               # (N,K,d)
    quad = (XL * X2d.unsqueeze(1)).sum(-1)
    lin = X2d @ niw_EinvSigmamu(s).T
    out = -0.5 * quad + lin - 0.5 * niw_EXTinvUX(s)
    return out + 0.5 * wishart_ElogdetinvSigma(s["invU"]) - 0.5 * s["dim"] * math.log(2 * math.pi)


def niw_raw_stats_exact(s, X, p):
    """dists/NormalInverseWishart.py:74-84: N = sum p, SExx = sum p x x^T, SEx = sum p x."""
    sample_shape = X.shape[:-s["event_dim"] - s["batch_dim"]]
    sd = tuple(range(len(sample_shape)))
    if p is None:
        SEx = X.sum(sd)
        SExx = (X.unsqueeze(-1) * X.unsqueeze(-2)).sum(sd)
        N = torch.tensor(float(math.prod(sample_shape)), dtype=X.dtype).expand(
            s["batch_shape"] + s["event_shape"][:-1])
    else:
        N = p.sum(sd)
        N = N.view(N.shape + (1,) * (s["event_dim"] - 1))
        pv = p.view(p.shape + (1,) * s["event_dim"])
        SExx = (X.unsqueeze(-1) * X.unsqueeze(-2) * pv.unsqueeze(-1)).sum(sd)
        SEx = (X * pv).sum(sd)
    return SExx, SEx, N


def weighted_gram_fast(Z2d, P2d, chunk=16384):
    """G_k = sum_n p_nk [z;1][z;1]^T for Z2d (N,D), P2d (N,K) -> (K,D+1,D+1), accumulated in fp64.
    Blocks of G are the reference's SExx/SEx/N (NIW :80-84) and SExx/SEyx/SEyy/SEx/SEy/N (MNW :185-202)."""
    N, D = Z2d.shape
    K = P2d.shape[1]
    G = torch.zeros(K, D + 1, D + 1, dtype=torch.float64)
    for a in range(0, N, chunk):
        z = Z2d[a:a + chunk].double()
        z1 = torch.cat([z, torch.ones(z.shape[0], 1, dtype=torch.float64)], -1)
        G += torch.einsum("nk,ni,nj->kij", P2d[a:a + chunk].double(), z1, z1)
    return G


def niw_ss_update(s, SExx, SEx, N, lr=1.0, beta=0.0):
    """dists/NormalInverseWishart.py:49-68."""
    if beta is not None:
        s["SExx"] = beta * s["SExx"] + SExx
        s["SEx"] = beta * s["SEx"] + SEx
        s["N"] = beta * s["N"] + N
        SExx, SEx, N = s["SExx"], s["SEx"], s["N"]
    lam0, mu0 = s["lambda_mu_0"], s["mu_0"]
    lam = lam0 + N
    mu = (lam0.unsqueeze(-1) * mu0 + SEx) / lam.unsqueeze(-1)
    S = SExx + lam0.unsqueeze(-1).unsqueeze(-1) * mu0.unsqueeze(-1) * mu0.unsqueeze(-2) \
        - lam.unsqueeze(-1).unsqueeze(-1) * mu.unsqueeze(-1) * mu.unsqueeze(-2)
    s["lambda_mu"] = lr * lam + (1 - lr) * s["lambda_mu"]
    s["mu"] = lr * mu + (1 - lr) * s["mu"]
    if s["fixed_precision"] is False:
        wishart_ss_update(s["invU"], S, N, lr)


def niw_raw_update_exact(s, X, p=None, lr=1.0, beta=None):
    """dists/NormalInverseWishart.py:70-86."""
    niw_ss_update(s, *niw_raw_stats_exact(s, X, p), lr=lr, beta=beta)


def niw_kl(s):
    """dists/NormalInverseWishart.py:99-105."""
    lam0, lam, d = s["lambda_mu_0"], s["lambda_mu"], s["dim"]
    KL = 0.5 * (lam0 / lam - 1 + (lam / lam0).log()) * d
    dm = s["mu"] - s["mu_0"]
    KL = KL + 0.5 * lam0 * (dm.unsqueeze(-1) * dm.unsqueeze(-2) * wishart_EinvSigma(s["invU"])).sum(-1).sum(-1)
    for _ in range(s["event_dim"] - 1):
        KL = KL.sum(-1)
    return KL + wishart_kl(s["invU"], extra_event_dims=s["event_dim"] - 1)


# --------------------------------------------------------------------------------------
# Gamma / NormalGamma (diagonal precision)   reference: dists/Gamma.py, dists/NormalGamma.py
# --------------------------------------------------------------------------------------

def gamma_new(event_shape, batch_shape=(), alpha0=1.0, beta0=1.0, dtype=torch.float32):
    """dists/Gamma.py:7-23: alpha = alpha_0 + rand, beta = beta_0 + rand (RNG, in this order)."""
    shape = tuple(batch_shape) + tuple(event_shape)
    a0 = torch.as_tensor(alpha0, dtype=dtype).expand(shape)
    b0 = torch.as_tensor(beta0, dtype=dtype).expand(shape)
    return {"event_dim": len(event_shape), "batch_dim": len(batch_shape), "alpha_0": a0, "beta_0": b0,
            "alpha": a0 + torch.rand(shape, dtype=dtype), "beta": b0 + torch.rand(shape, dtype=dtype),
            "SEx": 0.0, "SElogx": 0.0}


def gamma_ss_update(g, SElogx, SEx, lr=1.0, beta=None):
    """dists/Gamma.py:34-47."""
    if beta is not None:
        g["SEx"] = beta * g["SEx"] + SEx
        g["SElogx"] = beta * g["SElogx"] + SElogx
        SEx, SElogx = g["SEx"], g["SElogx"]
    g["alpha"] = (g["alpha_0"] + SElogx) * lr + g["alpha"] * (1 - lr)
    g["beta"] = (g["beta_0"] + SEx) * lr + g["beta"] * (1 - lr)


def gamma_mean(g):
    """dists/Gamma.py:93-94."""
    return g["alpha"] / g["beta"]


def gamma_meaninv(g):
    """dists/Gamma.py:99-100."""
    return g["beta"] / (g["alpha"] - 1)


def gamma_loggeomean(g):
    """dists/Gamma.py:105-106 (log alpha - log beta)."""
    return g["alpha"].log() - g["beta"].log()


def gamma_kl(g):
    """dists/Gamma.py:117-119."""
    a, b, a0, b0 = g["alpha"], g["beta"], g["alpha_0"], g["beta_0"]
    KL = (a - a0) * a.digamma() - a.lgamma() + a0.lgamma() + a0 * (b.log() - b0.log()) + a * (b0 / b - 1)
    return KL.sum(list(range(-g["event_dim"], 0)))


def ng_new(event_shape, batch_shape=(), scale=1.0, dtype=torch.float32):
    """dists/NormalGamma.py:6-28: lambda = lambda_0 + rand, Gamma(alpha_0 = 2, beta_0 = 2 scale^2), mu = mu_0 + randn /
    sqrt(gamma.mean()); event_dim is 1 whatever the event shape (:13)."""
    event_shape, batch_shape = tuple(event_shape), tuple(batch_shape)
    lam0 = torch.tensor(1.0, dtype=dtype).expand(batch_shape + event_shape[:-1])
    lam = lam0 + torch.rand_like(lam0)
    mu0 = torch.tensor(0.0, dtype=dtype).expand(batch_shape + event_shape)
    g = gamma_new(event_shape, batch_shape, 2.0, 2.0 * scale ** 2, dtype=dtype)
    mu = mu0 + torch.randn_like(mu0) / gamma_mean(g).sqrt()
    return {"kind": "ng", "dim": event_shape[-1], "event_shape": event_shape, "batch_shape": batch_shape,
            "event_dim": 1, "batch_dim": len(batch_shape), "lambda_mu_0": lam0, "lambda_mu": lam, "mu_0": mu0, "mu": mu,
            "gamma": g, "SExx": 0.0, "SEx": 0.0, "N": 0.0}


def ng_elog_like(s, X):
    """dists/NormalGamma.py:76-86: the expression that survives is :83,
    -1/2 ((X - mu)^2 gamma.mean()).sum(-1) + 1/2 gamma.loggeomean().sum(-1)."""
    out = -0.5 * ((X - s["mu"]) ** 2 * gamma_mean(s["gamma"])).sum(-1) + 0.5 * gamma_loggeomean(s["gamma"]).sum(-1)
    for _ in range(s["event_dim"] - 1):
        out = out.sum(-1)
    return out


def ng_raw_stats(s, X, p):
    """dists/NormalGamma.py:58-73."""
    sample_shape = X.shape[:-s["event_dim"] - s["batch_dim"]]
    sd = list(range(len(sample_shape)))
    if p is None:
        SEx = X.sum(sd)
        SExx = (X ** 2).sum(sd)
        N = torch.tensor(float(math.prod(sample_shape)), dtype=X.dtype).expand(s["batch_shape"] + s["event_shape"][:-1])
    else:
        N = p.sum(sd)
        pv = p.view(p.shape + s["event_dim"] * (1,))
        SEx = (X * pv).sum(sd)
        SExx = (X ** 2 * pv).sum(sd)
    return SExx, SEx, N


def ng_ss_update(s, SExx, SEx, N, lr=1.0, beta=None):
    """dists/NormalGamma.py:41-56."""
    if beta is not None:
        s["SExx"] = SExx + beta * s["SExx"]
        s["SEx"] = SEx + beta * s["SEx"]
        s["N"] = N + beta * s["N"]
        SExx, SEx, N = s["SExx"], s["SEx"], s["N"]
    lam0, mu0 = s["lambda_mu_0"], s["mu_0"]
    lam = lam0 + N
    mu = (lam0.unsqueeze(-1) * mu0 + SEx) / lam.unsqueeze(-1)
    SExx = SExx + lam0.unsqueeze(-1) * mu0 ** 2 - lam.unsqueeze(-1) * mu ** 2
    s["lambda_mu"] = lr * lam + (1 - lr) * s["lambda_mu"]
    s["mu"] = lr * mu + (1 - lr) * s["mu"]
    gamma_ss_update(s["gamma"], 0.5 * N.unsqueeze(-1), 0.5 * SExx, lr, beta)


def ng_kl(s):
    """dists/NormalGamma.py:88-94."""
    lam0, lam = s["lambda_mu_0"], s["lambda_mu"]
    out = lam0 / 2.0 * ((s["mu"] - s["mu_0"]) ** 2 * gamma_mean(s["gamma"])).sum(-1)
    out = out + s["dim"] / 2.0 * (lam0 / lam - (lam0 / lam).log() - 1)
    for _ in range(s["event_dim"] - 1):
        out = out.sum(-1)
    return out + gamma_kl(s["gamma"]).sum(-1)


# --------------------------------------------------------------------------------------
# Mixture / GaussianMixtureModel    reference: dists/Mixture.py, models/GaussianMixtureModel.py
# --------------------------------------------------------------------------------------

def stable_logsumexp(x, dims):
    """dists/Mixture.py:110-127 (list-of-dims branch, keepdim=False)."""
    xmax = x
    for d in dims:
        xmax = xmax.max(dim=d, keepdim=True)[0]
    y = (x - xmax).exp().sum(dim=dims, keepdim=False).log()
    for d in dims:
        xmax = xmax.squeeze(d)
    return xmax + y


def mixture_new(dist, event_shape, dtype=torch.float32):
    """dists/Mixture.py:8-19."""
    event_shape = tuple(event_shape)
    bs = dist["batch_shape"][:-len(event_shape)]
    return {"dist": dist, "event_shape": event_shape, "event_dim": len(event_shape),
            "batch_shape": bs, "batch_dim": len(bs),
            "pi": dirichlet_new(event_shape, bs, dtype=dtype),
            "logZ": torch.tensor(-torch.inf), "ELBO_last": torch.tensor(-torch.inf)}


def gmm_new(nc, dim, dtype=torch.float32, isotropic=False):
    """models/GaussianMixtureModel.py:7-12, scale = nc^(-1/dim): NormalInverseWishart components, or NormalGamma ones
    (isotropic=True).  As in the reference the component node is constructed before the Dirichlet (RNG order)."""
    make = ng_new if isotropic else niw_new
    return mixture_new(make((dim,), (nc,), scale=1.0 / nc ** (1.0 / dim), dtype=dtype), (nc,), dtype=dtype)


def _mixture_view(m, X):
    d = m["dist"]
    return X.view(X.shape[:-d["event_dim"]] + m["event_dim"] * (1,) + d["event_shape"])


def mixture_elog_like(m, X, exact=True):
    """dists/Mixture.py:68-70."""
    d = m["dist"]
    if d.get("kind") == "ng":
        return ng_elog_like(d, _mixture_view(m, X)) + dirichlet_loggeomean(m["pi"])
    if exact:
        return niw_elog_like_exact(d, _mixture_view(m, X)) + dirichlet_loggeomean(m["pi"])
    return niw_elog_like_fast(d, X) + dirichlet_loggeomean(m["pi"])


def mixture_update_assignments(m, X, exact=True, chunk=None):
    """dists/Mixture.py:38-45.  ``chunk`` evaluates the per-sample part in slices of the
    first sample dim (per-sample results are chunk independent)."""
    ed = list(range(-m["event_dim"], 0))
    if chunk is None:
        log_p = mixture_elog_like(m, X, exact)
    else:
        log_p = torch.cat([mixture_elog_like(m, X[a:a + chunk], exact) for a in range(0, X.shape[0], chunk)], 0)
    logZ = stable_logsumexp(log_p, ed)
    m["p"] = (log_p - logZ.view(logZ.shape + m["event_dim"] * (1,))).exp()
    sd = list(range(m["p"].ndim - m["batch_dim"] - m["event_dim"]))
    m["NA"] = m["p"].sum(sd)
    m["logZ"] = logZ.sum(sd)
    m["log_p"] = log_p


def mixture_kl(m):
    """dists/Mixture.py:72-73."""
    kl = ng_kl(m["dist"]) if m["dist"].get("kind") == "ng" else niw_kl(m["dist"])
    return kl.sum(list(range(-m["event_dim"], 0))) + dirichlet_kl(m["pi"])


def mixture_elbo(m):
    """dists/Mixture.py:75-76."""
    return m["logZ"] - mixture_kl(m)


def mixture_update_parms(m, X, lr=1.0, exact=True):
    """dists/Mixture.py:47-49, 65-66."""
    dirichlet_ss_update(m["pi"], m["NA"], lr=lr)
    d = m["dist"]
    if d.get("kind") == "ng":
        ng_ss_update(d, *ng_raw_stats(d, _mixture_view(m, X), m["p"]), lr=lr, beta=None)
    elif exact:
        niw_raw_update_exact(d, _mixture_view(m, X), m["p"], lr)
    else:
        G = weighted_gram_fast(X, m["p"]).to(X.dtype)
        D = d["dim"]
        niw_ss_update(d, G[:, :D, :D], G[:, :D, D], G[:, D, D], lr=lr, beta=None)


def mixture_update(m, X, iters=1, lr=1.0, exact=True, chunk=None):
    """dists/Mixture.py:54-62: E-step, ELBO (pre-M-step parameters), M-step.  Returns ELBO trace."""
    trace = []
    for _ in range(iters):
        mixture_update_assignments(m, X, exact, chunk)
        elbo = mixture_elbo(m)
        mixture_update_parms(m, X, lr, exact)
        m["ELBO_last"] = elbo
        trace.append(elbo)
    return trace


# --------------------------------------------------------------------------------------
# MatrixNormalWishart                reference: transforms/MatrixNormalWishart.py
# (no-mask branches only: SURVEY.md §2.1 #5 puts mask / X_mask out of scope)
# --------------------------------------------------------------------------------------

def mnw_new(event_shape, batch_shape=(), scale=1.0, pad_X=False, fixed_precision=False, dtype=torch.float32):
    """transforms/MatrixNormalWishart.py:20-70: mu = randn/sqrt(p') + mu_0, invV_0 = I, Wishart(n,n)."""
    event_shape, batch_shape = tuple(event_shape), tuple(batch_shape)
    n, p = event_shape[-2], event_shape[-1]
    if pad_X:
        p = p + 1
        event_shape = event_shape[:-1] + (p,)
    mu0 = torch.tensor(0.0, dtype=dtype).expand(batch_shape + event_shape)
    mu = torch.randn_like(mu0) / torch.sqrt(torch.tensor(p, dtype=dtype)) + mu0
    invV0 = torch.eye(p, dtype=dtype).expand(batch_shape + event_shape[:-2] + (p, p))
    return {
        "n": n, "p": p, "pad_X": pad_X, "fixed_precision": fixed_precision,
        "event_shape": event_shape, "event_dim": len(event_shape),
        "batch_shape": batch_shape, "batch_dim": len(batch_shape),
        "mu_0": mu0, "mu": mu, "invV_0": invV0, "invV": invV0, "V": invV0.inverse(),
        "logdetinvV": invV0.logdet(), "logdetinvV_0": invV0.logdet(),
        "invU": wishart_new(n, batch_shape, scale, extra_event=event_shape[:-2], dtype=dtype),
        "SEyy": 0.0, "SExx": 0.0, "SEyx": 0.0, "N": 0.0,
        "log2pi": math.log(2 * math.pi),
    }


def mnw_EinvUX(s):
    """transforms/MatrixNormalWishart.py:419-420."""
    return wishart_EinvSigma(s["invU"]) @ s["mu"]


def mnw_EXTinvUX(s):
    """transforms/MatrixNormalWishart.py:437-438."""
    return s["n"] * s["V"] + s["mu"].transpose(-1, -2) @ wishart_EinvSigma(s["invU"]) @ s["mu"]


def mnw_elog_like_exact(s, X, Y):
    """transforms/MatrixNormalWishart.py:219-232 (X: (...,p,1), Y: (...,n,1))."""
    ES = wishart_EinvSigma(s["invU"])
    ELL = -0.5 * (Y.transpose(-2, -1) @ ES @ Y).squeeze(-1).squeeze(-1)
    A, B = mnw_EinvUX(s), mnw_EXTinvUX(s)
    if s["pad_X"]:
        ELL = ELL + (Y.transpose(-2, -1) @ (A[..., :, :-1] @ X + A[..., :, -1:])).squeeze(-1).squeeze(-1)
        ELL = ELL - 0.5 * (X.transpose(-2, -1) @ B[..., :-1, :-1] @ X + 2 * B[..., -1:, :-1] @ X
                           + B[..., -1:, -1:]).squeeze(-1).squeeze(-1)
    else:
        ELL = ELL + (Y.transpose(-2, -1) @ A @ X).squeeze(-1).squeeze(-1)
        ELL = ELL - 0.5 * (X.transpose(-2, -1) @ B @ X).squeeze(-1).squeeze(-1)
    ELL = ELL + 0.5 * wishart_ElogdetinvSigma(s["invU"]) - 0.5 * s["n"] * s["log2pi"]
    for _ in range(s["event_dim"] - 2):
        ELL = ELL.sum(-1)
    return ELL


def mnw_elog_like_fast(s, X2d, Y2d):
    """Same quantity for batch (K,), event (n,p'), X2d (N,p), Y2d (N,n): residual form
    -1/2 nu (y - mu x~)^T U (y - mu x~) - 1/2 n x~^T V x~ + const (SURVEY.md §3.2 identity)."""
    N = X2d.shape[0]
    Xt = torch.cat([X2d, torch.ones(N, 1, dtype=X2d.dtype)], -1) if s["pad_X"] else X2d
    ES = wishart_EinvSigma(s["invU"])                          # (K,n,n)
    R = Y2d.unsqueeze(1) - torch.einsum("kij,nj->nki", s["mu"], Xt)       # (N,K,n)
    q1 = torch.einsum("nki,kij,nkj->nk", R, ES, R)
    q2 = s["n"] * torch.einsum("ni,kij,nj->nk", Xt, s["V"], Xt)
    return -0.5 * q1 - 0.5 * q2 + 0.5 * wishart_ElogdetinvSigma(s["invU"]) - 0.5 * s["n"] * s["log2pi"]


def mnw_raw_stats_exact(s, X, Y, p):
    """transforms/MatrixNormalWishart.py:174-202 incl. the pad_X concatenations."""
    sample_shape = X.shape[:-s["event_dim"] - s["batch_dim"]]
    sd = tuple(range(len(sample_shape)))
    if p is None:
        SExx = (X * X.transpose(-2, -1)).sum(sd)
        SEyy = (Y * Y.transpose(-2, -1)).sum(sd)
        SEyx = (Y * X.transpose(-2, -1)).sum(sd)
        N = torch.tensor(float(math.prod(sample_shape)), dtype=X.dtype).expand(
            s["batch_shape"] + s["event_shape"][:-2])
        pv = None
    else:
        N = p.sum(sd)
        pv = p.view(p.shape + s["event_dim"] * (1,))
        SExx = (X * X.transpose(-2, -1) * pv).sum(sd)
        SEyy = (Y * Y.transpose(-2, -1) * pv).sum(sd)
        SEyx = (Y * X.transpose(-2, -1) * pv).sum(sd)
    if s["pad_X"]:
        SEx = X.sum(sd) if pv is None else (X * pv).sum(sd)
        SEy = Y.sum(sd) if pv is None else (Y * pv).sum(sd)
        SExx = torch.cat((SExx, SEx), dim=-1)
        SEx1 = torch.cat((SEx, N.view(N.shape + (1, 1))), dim=-2)
        SExx = torch.cat((SExx, SEx1.transpose(-2, -1)), dim=-2)
        SEyx = torch.cat((SEyx, SEy.expand(SEyx.shape[:-1] + (1,))), dim=-1)
    return SExx, SEyx, SEyy, N


def mnw_ss_update(s, SExx, SEyx, SEyy, N, lr=1.0, beta=None):
    """transforms/MatrixNormalWishart.py:82-141, no-mask branch (:105-108, :122-135)."""
    if beta is not None:
        s["SExx"] = beta * s["SExx"] + SExx
        s["SEyx"] = beta * s["SEyx"] + SEyx
        s["SEyy"] = beta * s["SEyy"] + SEyy
        s["N"] = beta * s["N"] + N
        SExx, SEyx, SEyy, N = s["SExx"], s["SEyx"], s["SEyy"], s["N"]
    invV = s["invV_0"] + SExx
    muinvV = s["mu_0"] @ s["invV_0"] + SEyx
    mu = torch.linalg.solve(invV, muinvV.transpose(-2, -1)).transpose(-2, -1)
    if s["fixed_precision"] is False:
        SEyy = SEyy - mu @ invV @ mu.transpose(-2, -1) + s["mu_0"] @ s["invV_0"] @ s["mu_0"].transpose(-2, -1)
        wishart_ss_update(s["invU"], SEyy, N, lr=lr, beta=None)
    s["invV"] = lr * invV + (1.0 - lr) * s["invV"]
    s["invV"] = 0.5 * (s["invV"] + s["invV"].transpose(-2, -1))
    s["mu"] = lr * mu + (1.0 - lr) * s["mu"]
    s["V"] = s["invV"].inverse()
    s["logdetinvV"] = s["invV"].logdet()


def mnw_gram_blocks(s, G):
    """Blocks of G = sum p z~ z~^T, z = [y; x] (+1), laid out as mnw_ss_update expects."""
    n, pp = s["n"], s["p"]
    p = pp - 1 if s["pad_X"] else pp
    D = n + p
    SEyy = G[..., :n, :n]
    if s["pad_X"]:
        idx = list(range(n, D + 1))
        SExx = G[..., idx, :][..., :, idx]
        SEyx = G[..., :n, :][..., :, idx]
    else:
        SExx = G[..., n:D, n:D]
        SEyx = G[..., :n, n:D]
    return SExx, SEyx, SEyy, G[..., D, D]


def mnw_kl(s):
    """transforms/MatrixNormalWishart.py:206-216 (X_mask None)."""
    n, p = s["n"], s["p"]
    KL = n / 2.0 * s["logdetinvV"] - n / 2.0 * s["logdetinvV_0"] - n * p / 2.0
    KL = KL + 0.5 * n * (s["invV_0"] * s["V"]).sum(-1).sum(-1)
    dm = s["mu"] - s["mu_0"]
    temp = dm.transpose(-2, -1) @ wishart_EinvSigma(s["invU"]) @ dm
    KL = KL + 0.5 * (s["invV_0"] * temp).sum(-1).sum(-1)
    for _ in range(s["event_dim"] - 2):
        KL = KL.sum(-1)
    return KL + wishart_kl(s["invU"], extra_event_dims=s["event_dim"] - 2)


# --------------------------------------------------------------------------------------
# MatrixNormalGamma (diagonal output precision)   reference: transforms/MatrixNormalGamma.py, dists/DiagonalWishart.py
# --------------------------------------------------------------------------------------

def mng_new(event_shape, batch_shape=(), scale=1.0, pad_X=False, fixed_precision=False, dtype=torch.float32):
    """transforms/MatrixNormalGamma.py:22-86 (no masks): mu = randn / sqrt(p') (no mu_0 term), invV_0 = I,
    DiagonalWishart -> Gamma(alpha_0 = 2, beta_0 = scale^2 / 0.5) over the n outputs (dists/DiagonalWishart.py:9-20)."""
    event_shape, batch_shape = tuple(event_shape), tuple(batch_shape)
    n, p = event_shape[-2], event_shape[-1]
    if pad_X:
        p = p + 1
        event_shape = event_shape[:-1] + (p,)
    mu0 = torch.tensor(0.0, dtype=dtype).expand(batch_shape + event_shape)
    mu = torch.randn_like(mu0) / torch.sqrt(torch.tensor(float(p), dtype=dtype))
    invV0 = torch.eye(p, dtype=dtype).expand(batch_shape + event_shape[:-2] + (p, p))
    return {
        "kind": "mng", "n": n, "p": p, "pad_X": pad_X, "fixed_precision": fixed_precision,
        "event_shape": event_shape, "event_dim": len(event_shape), "batch_shape": batch_shape, "batch_dim": len(batch_shape),
        "mu_0": mu0, "mu": mu, "invV_0": invV0, "invV": invV0, "V": invV0.inverse(),
        "logdetinvV": invV0.logdet(), "logdetinvV_0": invV0.logdet(),
        "invU": {"gamma": gamma_new(event_shape[:-1], batch_shape, 2.0, scale ** 2 / 0.5, dtype=dtype)},
        "SEyy": 0.0, "SExx": 0.0, "SEyx": 0.0, "N": 0.0, "log2pi": math.log(2 * math.pi),
    }


def _diag_embed(v):
    return v.unsqueeze(-1) * torch.eye(v.shape[-1], dtype=v.dtype)


def mng_EinvSigma(s):
    """transforms/MatrixNormalGamma.py:466-467 -> dists/DiagonalWishart.py:56-57."""
    return _diag_embed(gamma_mean(s["invU"]["gamma"]))


def mng_EinvUX(s):
    """transforms/MatrixNormalGamma.py:421-422."""
    return gamma_mean(s["invU"]["gamma"]).unsqueeze(-1) * s["mu"]


def mng_EXTinvUX(s):
    """transforms/MatrixNormalGamma.py:439-440."""
    return s["n"] * s["V"] + s["mu"].transpose(-1, -2) @ (gamma_mean(s["invU"]["gamma"]).unsqueeze(-1) * s["mu"])


def mng_elog_like(s, X, Y):
    """transforms/MatrixNormalGamma.py:227-243 (X: (...,p,1), Y: (...,n,1))."""
    ELL = -0.5 * (Y.transpose(-2, -1) @ mng_EinvSigma(s) @ Y).squeeze(-1).squeeze(-1)
    A, B = mng_EinvUX(s), mng_EXTinvUX(s)
    if s["pad_X"]:
        ELL = ELL + (Y.transpose(-2, -1) @ (A[..., :, :-1] @ X + A[..., :, -1:])).squeeze(-1).squeeze(-1)
        ELL = ELL - 0.5 * (X.transpose(-2, -1) @ B[..., :-1, :-1] @ X + 2 * B[..., -1:, :-1] @ X
                           + B[..., -1:, -1:]).squeeze(-1).squeeze(-1)
    else:
        ELL = ELL + (Y.transpose(-2, -1) @ A @ X).squeeze(-1).squeeze(-1)
        ELL = ELL - 0.5 * (X.transpose(-2, -1) @ B @ X).squeeze(-1).squeeze(-1)
    ELL = ELL + 0.5 * gamma_loggeomean(s["invU"]["gamma"]).sum(-1) - 0.5 * s["n"] * s["log2pi"]
    for _ in range(s["event_dim"] - 2):
        ELL = ELL.sum(-1)
    return ELL


def mng_ss_update(s, SExx, SEyx, SEyy, N, lr=1.0, beta=None):
    """transforms/MatrixNormalGamma.py:87-141, no-mask branch (:112-116, :129-141)."""
    if beta is not None:
        s["SExx"] = beta * s["SExx"] + SExx
        s["SEyx"] = beta * s["SEyx"] + SEyx
        s["SEyy"] = beta * s["SEyy"] + SEyy
        s["N"] = beta * s["N"] + N
        SExx, SEyx, SEyy, N = s["SExx"], s["SEyx"], s["SEyy"], s["N"]
    invV = s["invV_0"] + SExx
    muinvV = s["mu_0"] @ s["invV_0"] + SEyx
    mu = torch.linalg.solve(invV, muinvV.transpose(-2, -1)).transpose(-2, -1)
    if s["fixed_precision"] is False:
        SEyy = SEyy - mu @ invV @ mu.transpose(-2, -1) + s["mu_0"] @ s["invV_0"] @ s["mu_0"].transpose(-2, -1)
        # dists/DiagonalWishart.py:32-37: gamma.ss_update(N / 2, diag / 2, lr, beta=None)
        gamma_ss_update(s["invU"]["gamma"], N.unsqueeze(-1) / 2.0, SEyy.diagonal(dim1=-2, dim2=-1) / 2.0, lr, None)
    s["invV"] = lr * invV + (1.0 - lr) * s["invV"]
    s["invV"] = 0.5 * (s["invV"] + s["invV"].transpose(-2, -1))
    s["mu"] = lr * mu + (1.0 - lr) * s["mu"]
    s["V"] = s["invV"].inverse()
    s["logdetinvV"] = s["invV"].logdet()


def mng_kl(s):
    """transforms/MatrixNormalGamma.py:206-225 (X_mask None, uniform_precision False)."""
    n, p = s["n"], s["p"]
    KL = n / 2.0 * s["logdetinvV"] - n / 2.0 * s["logdetinvV_0"] - n * p / 2.0
    KL = KL + 0.5 * n * (s["invV_0"] * s["V"]).sum(-1).sum(-1)
    dm = s["mu"] - s["mu_0"]
    temp = dm.transpose(-2, -1) @ (gamma_mean(s["invU"]["gamma"]).unsqueeze(-1) * dm)
    KL = KL + 0.5 * (s["invV_0"] * temp).sum(-1).sum(-1)
    for _ in range(s["event_dim"] - 2):
        KL = KL.sum(-1)
    KL = KL + gamma_kl(s["invU"]["gamma"])
    for _ in range(s["event_dim"] - 2):
        KL = KL.sum(-1)
    return KL


# --------------------------------------------------------------------------------------
# MixtureofLinearTransforms           reference: transforms/MixtureofLinearTransforms.py
# --------------------------------------------------------------------------------------

def molt_new(n, p, dim, pad_X=True, dtype=torch.float32, type='Wishart'):
    """transforms/MixtureofLinearTransforms.py:12-32 (batch_shape=()): MatrixNormalWishart experts, or MatrixNormalGamma
    ones (type='Gamma')."""
    make = mng_new if type == 'Gamma' else mnw_new
    W = make((n, p), (dim,), scale=1.0 / dim ** (1.0 / n), pad_X=pad_X, dtype=dtype)
    return {"n": n, "p": p, "dim": dim, "W": W, "pi": dirichlet_new((dim,), dtype=dtype),
            "ELBO_last": -torch.tensor(torch.inf)}


def molt_update_assignments(m, X, Y, exact=True, chunk=None):
    """transforms/MixtureofLinearTransforms.py:34-41: max-shift softmax, per-sample logZ."""
    def ell(Xc, Yc):
        if m["W"].get("kind") == "mng":
            return mng_elog_like(m["W"], Xc.unsqueeze(-3), Yc.unsqueeze(-3))
        if exact:
            return mnw_elog_like_exact(m["W"], Xc.unsqueeze(-3), Yc.unsqueeze(-3))
        return mnw_elog_like_fast(m["W"], Xc.squeeze(-1), Yc.squeeze(-1))
    if chunk is None:
        log_p = ell(X, Y)
    else:
        log_p = torch.cat([ell(X[a:a + chunk], Y[a:a + chunk]) for a in range(0, X.shape[0], chunk)], 0)
    log_p = log_p + dirichlet_loggeomean(m["pi"])
    m["log_p"] = log_p
    shift = log_p.max(-1, True)[0]
    pu = (log_p - shift).exp()
    Z = pu.sum(-1, True)
    m["p"] = pu / Z
    m["logZ"] = (Z.log() + shift).squeeze(-1)


def molt_kl(m):
    """transforms/MixtureofLinearTransforms.py:123-124."""
    return dirichlet_kl(m["pi"]) + (mng_kl(m["W"]) if m["W"].get("kind") == "mng" else mnw_kl(m["W"])).sum(-1)


def molt_elbo(m):
    """transforms/MixtureofLinearTransforms.py:126-130 (batch_dim = 0)."""
    return m["logZ"].sum() - molt_kl(m)


def molt_raw_update(m, X, Y, iters=1, lr=1.0, exact=True, chunk=None):
    """transforms/MixtureofLinearTransforms.py:50-61.  Returns the ELBO trace."""
    trace = []
    for _ in range(iters):
        molt_update_assignments(m, X, Y, exact, chunk)
        elbo = molt_elbo(m)
        dirichlet_ss_update(m["pi"], m["p"].sum(0), lr=lr)
        W = m["W"]
        if W.get("kind") == "mng":     # transforms/MatrixNormalGamma.py:174-204 forms the same statistics as the Wishart node
            mng_ss_update(W, *mnw_raw_stats_exact(W, X.unsqueeze(-3), Y.unsqueeze(-3), m["p"]), lr=lr, beta=None)
        elif exact:
            mnw_ss_update(W, *mnw_raw_stats_exact(W, X.unsqueeze(-3), Y.unsqueeze(-3), m["p"]), lr=lr, beta=None)
        else:
            Z = torch.cat([Y.squeeze(-1), X.squeeze(-1)], -1)
            G = weighted_gram_fast(Z, m["p"]).to(X.dtype)
            mnw_ss_update(W, *mnw_gram_blocks(W, G), lr=lr, beta=None)
        m["ELBO_last"] = elbo
        trace.append(elbo)
    return trace


def mnw_elog_like_given(s, mux, EXXTx, muy, EXXTy):
    """transforms/MatrixNormalWishart.py:234-249 (Elog_like_given_pX_pY): expected log likelihood under Gaussian beliefs
    about input and output, given their means (..., p, 1) / (..., n, 1) and second moments E[xx^T], E[yy^T]."""
    ES, EinvUX, EXTinvUX = wishart_EinvSigma(s["invU"]), mnw_EinvUX(s), mnw_EXTinvUX(s)
    ELL = -0.5 * (EXXTy * ES).sum(-1).sum(-1)
    if s["pad_X"]:
        ELL = ELL + (muy.transpose(-2, -1) @ (EinvUX[..., :, :-1] @ mux + EinvUX[..., :, -1:])).squeeze(-1).squeeze(-1)
        ELL = ELL - 0.5 * (EXXTx * EXTinvUX[..., :-1, :-1]).sum(-1).sum(-1)
        ELL = ELL - (EXTinvUX[..., -1:, :-1] @ mux).squeeze(-1).squeeze(-1)
        ELL = ELL - 0.5 * (EXTinvUX[..., -1, -1])
    else:
        ELL = ELL + (muy.transpose(-2, -1) @ EinvUX @ mux).squeeze(-1).squeeze(-1)
        ELL = ELL - 0.5 * (EXXTx * EXTinvUX).sum(-1).sum(-1)
    return ELL + 0.5 * wishart_ElogdetinvSigma(s["invU"]) - 0.5 * s["n"] * torch.log(2 * torch.tensor(torch.pi, dtype=mux.dtype))


def mnw_stats_given(s, mux, EXXTx, muy, EXXTy, p):
    """transforms/MatrixNormalWishart.py:143-170: the statistics of update(pX, pY, p) (no cross-covariance between the two
    beliefs: SEyx uses the means only)."""
    sample_shape = mux.shape[:-s["event_dim"] - s["batch_dim"]]
    sd = tuple(range(len(sample_shape)))
    if p is None:
        N = torch.tensor(float(math.prod(sample_shape)), dtype=mux.dtype).expand(s["batch_shape"] + s["event_shape"][:-2])
        w = lambda t: t.sum(sd)                                                  # noqa: E731
    else:
        N = p.sum(sd)
        pv = p.view(p.shape + s["event_dim"] * (1,))
        w = lambda t: (t * pv).sum(sd)                                           # noqa: E731
    SExx, SEyy, SEyx = w(EXXTx), w(EXXTy), w(muy @ mux.transpose(-2, -1))
    if s["pad_X"]:
        SEx, SEy = w(mux), w(muy)
        SExx = torch.cat((SExx, SEx), dim=-1)
        SEx1 = torch.cat((SEx, N.view(N.shape + (1, 1))), dim=-2)
        SExx = torch.cat((SExx, SEx1.transpose(-2, -1)), dim=-2)
        SEyx = torch.cat((SEyx, SEy.expand(SEyx.shape[:-1] + (1,))), dim=-1)
    return SExx, SEyx, SEyy, N


def molt_update_given(m, mux, Sx, muy, Sy, lr=1.0):
    """transforms/MixtureofLinearTransforms.py:62-69, 77-90: one E + M iteration of update(pX, pY) for beliefs with means
    mux (N,p,1), muy (N,n,1) and covariances Sx, Sy.  Returns the ELBO (computed between E and M)."""
    mx, my = mux.unsqueeze(-3), muy.unsqueeze(-3)
    Exx = (Sx + mux @ mux.transpose(-2, -1)).unsqueeze(-3)
    Eyy = (Sy + muy @ muy.transpose(-2, -1)).unsqueeze(-3)
    log_p = mnw_elog_like_given(m["W"], mx, Exx, my, Eyy) + dirichlet_loggeomean(m["pi"])
    shift = log_p.max(-1, True)[0]
    pu = (log_p - shift).exp()
    Z = pu.sum(-1, True)
    m["p"] = pu / Z
    m["logZ"] = (Z.log() + shift).squeeze(-1)
    m["log_p"] = log_p
    elbo = molt_elbo(m)
    dirichlet_ss_update(m["pi"], m["p"].sum(0), lr=lr)
    mnw_ss_update(m["W"], *mnw_stats_given(m["W"], mx, Exx, my, Eyy, m["p"]), lr=lr, beta=None)
    m["ELBO_last"] = elbo
    return elbo


def mnw_predict(s, X):
    """transforms/MatrixNormalWishart.py:381-390 (pad_X branch :383-384): natural parameters of p(y | x) per component and
    the per-component log evidence Res - pY.Res() (dists/MultivariateNormal_vector_format.py:118-119).  X: (..., p, 1)."""
    EinvUX, EXTinvUX = mnw_EinvUX(s), mnw_EXTinvUX(s)
    n = s["n"]
    log2pi = torch.log(2 * torch.tensor(torch.pi, dtype=X.dtype))
    if s["pad_X"]:
        invSigmamu_y = EinvUX[..., :, :-1] @ X + EinvUX[..., :, -1:]
        Res = -0.5 * X.transpose(-1, -2) @ EXTinvUX[..., :-1, :-1] @ X - EXTinvUX[..., -1:, :-1] @ X - 0.5 * EXTinvUX[..., -1:, -1:]
    else:
        invSigmamu_y = EinvUX @ X
        Res = -0.5 * X.transpose(-1, -2) @ EXTinvUX @ X
    Res = Res.squeeze(-1).squeeze(-1) + 0.5 * wishart_ElogdetinvSigma(s["invU"]) - 0.5 * n * log2pi
    invSigma = wishart_EinvSigma(s["invU"])
    mean = invSigma.inverse() @ invSigmamu_y
    pY_Res = -0.5 * (mean * invSigmamu_y).sum(-1).sum(-1) + 0.5 * invSigma.logdet() - 0.5 * n * log2pi
    return invSigma, invSigmamu_y, mean, Res - pY_Res


def molt_predict(m, X):
    """transforms/MixtureofLinearTransforms.py:91-108: mixture-of-experts predictive mean / covariance and the gate
    probabilities for inputs X (N, p, 1).  Returns (mu (N,n,1), Sigma (N,n,n), p (N,K))."""
    invSigma, _, mean, Res = mnw_predict(m["W"], X.unsqueeze(-3))
    log_p = Res + dirichlet_loggeomean(m["pi"])
    log_p = log_p - log_p.max(-1, True)[0]
    p = log_p.exp()
    p = p / p.sum(-1, True)
    pe = p.unsqueeze(-1).unsqueeze(-1)
    Sigma = ((invSigma.inverse() + mean @ mean.transpose(-2, -1)) * pe).sum(-3)
    mu = (mean * pe).sum(-3)
    Sigma = Sigma - mu @ mu.transpose(-2, -1)
    return mu, Sigma, p


# --------------------------------------------------------------------------------------
# HMM / ARHMM                          reference: models/HMM.py, models/ARHMM.py
# --------------------------------------------------------------------------------------

def _lse(x, dim, keepdim=False):
    """utils/torch_functions.py:2-4 (stable_logsumexp)."""
    xmax = x.amax(dim=dim, keepdim=True)
    out = xmax + (x - xmax).exp().sum(dim=dim, keepdim=True).log()
    if keepdim:
        return out
    if isinstance(dim, int):
        return out.squeeze(dim)
    for d in sorted([dd % x.ndim for dd in dim], reverse=True):
        out = out.squeeze(d)
    return out


def hmm_new(obs, K, dtype=torch.float32):
    """models/HMM.py:6-31 (batch_shape = (), no transition mask): transition prior eye + 0.5."""
    alpha = torch.eye(K, dtype=dtype) + 0.5
    tr = {"event_dim": 1, "batch_dim": 1, "alpha_0": alpha,
          "alpha": alpha * (1.0 + torch.rand(alpha.shape, dtype=dtype)), "NA": 0.0}
    return {"obs": obs, "dim": K, "transition": tr, "initial": dirichlet_new((K,), dtype=dtype),
            "p": None, "ptemp": 1.0, "logZ": torch.tensor(-torch.inf), "ELBO_last": torch.tensor(-torch.inf)}


def hmm_forward_backward_logits(h, fw):
    """models/HMM.py:72-105 (log-space forward pass, backward smoothing, SEzz/SEz0)."""
    fw = fw.clone()
    tr = dirichlet_loggeomean(h["transition"])
    init = dirichlet_loggeomean(h["initial"])
    T = fw.shape[0]
    fw[0] = _lse(init.unsqueeze(-1) + tr + fw[0].unsqueeze(-2), -2)
    for t in range(1, T):
        fw[t] = _lse(fw[t - 1].unsqueeze(-1) + tr + fw[t].unsqueeze(-2), -2)
    logZ = _lse(fw[-1], -1, True)
    fw = fw - logZ
    logZ = logZ.squeeze(-1)
    SEzz = torch.zeros(fw.shape[1:] + (h["dim"],), dtype=fw.dtype)
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
    p = ((fw - fw.max(-1, keepdim=True)[0]) / h["ptemp"]).exp()
    p = p / p.sum(-1, keepdim=True)
    return p, SEzz, SEz0, logZ


def arhmm_new(K, n, p, pad_X=True, dtype=torch.float32):
    """models/ARHMM.py:14-16: MNW(event=(n,p), batch=(K,), pad_X) emissions under an HMM."""
    return hmm_new(mnw_new((n, p), (K,), pad_X=pad_X, dtype=dtype), K, dtype=dtype)


def arhmm_obs_logits(h, X, Y, exact=True):
    """models/ARHMM.py:18-22."""
    if exact:
        return mnw_elog_like_exact(h["obs"], X, Y)
    T, S = X.shape[:2]
    return mnw_elog_like_fast(h["obs"], X.reshape(T * S, -1), Y.reshape(T * S, -1)).view(T, S, -1)


def hmm_kl(h, obs_kl):
    """models/HMM.py:154-156."""
    return obs_kl.sum(-1) + dirichlet_kl(h["transition"]).sum(-1) + dirichlet_kl(h["initial"])


def arhmm_update(h, X, Y, iters=1, lr=1.0, beta=None, exact=True):
    """models/HMM.py:141-152 with models/ARHMM.py:24-25 (ELBO evaluated AFTER the M-step)."""
    trace = []
    for _ in range(iters):
        p, SEzz, SEz0, logZ = hmm_forward_backward_logits(h, arhmm_obs_logits(h, X, Y, exact))
        h["p"] = p
        NA = p.sum(0)
        sd = list(range(NA.ndim - 1))
        h["NA"], SEzz, SEz0, h["logZ"] = NA.sum(sd), SEzz.sum(sd), SEz0.sum(sd), logZ.sum(sd)
        dirichlet_ss_update(h["transition"], SEzz, lr=lr, beta=beta)
        dirichlet_ss_update(h["initial"], SEz0, lr=lr, beta=beta)
        W = h["obs"]
        if exact:
            mnw_ss_update(W, *mnw_raw_stats_exact(W, X, Y, p), lr=lr, beta=beta)
        else:
            T, S = X.shape[:2]
            Z = torch.cat([Y.reshape(T * S, -1), X.reshape(T * S, -1)], -1)
            G = weighted_gram_fast(Z, p.reshape(T * S, -1)).to(X.dtype)
            mnw_ss_update(W, *mnw_gram_blocks(W, G), lr=lr, beta=beta)
        elbo = h["logZ"] - hmm_kl(h, mnw_kl(W))
        h["ELBO_last"] = elbo
        trace.append(elbo)
    return trace


def hmm_niw_update(h, y, iters=1, lr=1.0, beta=None):
    """models/HMM.py:113-152 for a NormalInverseWishart emission node (batch_shape = (K,), any event_dim): obs_logits (:113-117),
    update_states (:119-132), update_markov_parms (:134-136), update_obs_parms (:138-139); ELBO evaluated AFTER the M-step."""
    trace = []
    s = h["obs"]
    for _ in range(iters):
        Xv = y.unsqueeze(-1 - s["event_dim"])
        p, SEzz, SEz0, logZ = hmm_forward_backward_logits(h, niw_elog_like_exact(s, Xv))
        h["p"] = p
        NA = p.sum(0)
        sd = list(range(NA.ndim - 1))
        h["NA"], SEzz, SEz0, h["logZ"] = NA.sum(sd), SEzz.sum(sd), SEz0.sum(sd), logZ.sum(sd)
        dirichlet_ss_update(h["transition"], SEzz, lr=lr, beta=beta)
        dirichlet_ss_update(h["initial"], SEz0, lr=lr, beta=beta)
        niw_raw_update_exact(s, Xv, p, lr=lr, beta=beta)
        elbo = h["logZ"] - hmm_kl(h, niw_kl(s))
        h["ELBO_last"] = elbo
        trace.append(elbo)
    return trace


def arhmm_prxy_update(h, mux, Sx, muy, Sy, iters=1, lr=1.0, beta=None):
    """models/HMM.py:141-152 with models/ARHMM.py:35-46 (ARHMM_prXY): observation logits from Elog_like_given_pX_pY, the
    observation update from update(pX, pY, p).  Beliefs: means (T,S,1,p,1) / (T,S,1,n,1), covariances (T,S,1,p,p) / (T,S,1,n,n)."""
    W = h["obs"]
    Exx = Sx + mux @ mux.transpose(-2, -1)
    Eyy = Sy + muy @ muy.transpose(-2, -1)
    trace = []
    for _ in range(iters):
        p, SEzz, SEz0, logZ = hmm_forward_backward_logits(h, mnw_elog_like_given(W, mux, Exx, muy, Eyy))
        h["p"] = p
        NA = p.sum(0)
        sd = list(range(NA.ndim - 1))
        h["NA"], SEzz, SEz0, h["logZ"] = NA.sum(sd), SEzz.sum(sd), SEz0.sum(sd), logZ.sum(sd)
        dirichlet_ss_update(h["transition"], SEzz, lr=lr, beta=beta)
        dirichlet_ss_update(h["initial"], SEz0, lr=lr, beta=beta)
        mnw_ss_update(W, *mnw_stats_given(W, mux, Exx, muy, Eyy, p), lr=lr, beta=beta)
        elbo = h["logZ"] - hmm_kl(h, mnw_kl(W))
        h["ELBO_last"] = elbo
        trace.append(elbo)
    return trace


def arhmm_prxry_new(K, n, p1, p2, pad_X=False, dtype=torch.float32):
    """models/ARHMM.py:56-60: MNW(event=(n, p1 + p2), batch=(K,), pad_X=False by default) emissions under an HMM."""
    return hmm_new(mnw_new((n, p1 + p2), (K,), pad_X=pad_X, dtype=dtype), K, dtype=dtype)


def arhmm_prxry_beliefs(mux, Sx, R, Y):
    """models/ARHMM.py:65-71: the stacked regressor belief — mean [E x; r], covariance diag(Sigma_x, 0) — and the observed
    outputs as a point mass.  Returns (mu, E[xx^T], y, yy^T) as mnw_elog_like_given / mnw_stats_given take them."""
    p1, p2 = mux.shape[-2], R.shape[-2]
    Sigma = torch.zeros(Sx.shape[:-2] + (p1 + p2, p1 + p2), dtype=Sx.dtype)
    Sigma[..., :p1, :p1] = Sx
    mu = torch.cat((mux, R), dim=-2)
    return mu, Sigma + mu @ mu.transpose(-2, -1), Y, Y @ Y.transpose(-2, -1)


def arhmm_prxry_update(h, mux, Sx, R, Y, iters=1, lr=1.0, beta=None):
    """models/HMM.py:141-152 with models/ARHMM.py:62-77 (ARHMM_prXRY).  Beliefs / data: mux (T,S,1,p1,1), Sx (T,S,1,p1,p1),
    R (T,S,1,p2,1), Y (T,S,1,n,1)."""
    W = h["obs"]
    mu, Exx, y, Eyy = arhmm_prxry_beliefs(mux, Sx, R, Y)
    trace = []
    for _ in range(iters):
        p, SEzz, SEz0, logZ = hmm_forward_backward_logits(h, mnw_elog_like_given(W, mu, Exx, y, Eyy))
        h["p"] = p
        NA = p.sum(0)
        sd = list(range(NA.ndim - 1))
        h["NA"], SEzz, SEz0, h["logZ"] = NA.sum(sd), SEzz.sum(sd), SEz0.sum(sd), logZ.sum(sd)
        dirichlet_ss_update(h["transition"], SEzz, lr=lr, beta=beta)
        dirichlet_ss_update(h["initial"], SEz0, lr=lr, beta=beta)
        mnw_ss_update(W, *mnw_stats_given(W, mu, Exx, y, Eyy, p), lr=lr, beta=beta)
        elbo = h["logZ"] - hmm_kl(h, mnw_kl(W))
        h["ELBO_last"] = elbo
        trace.append(elbo)
    return trace


# --------------------------------------------------------------------------------------
# state (de)serialisation helpers used by the golden fixtures and the GPU parity tests
# --------------------------------------------------------------------------------------

NIW_KEYS = ("lambda_mu_0", "lambda_mu", "mu_0", "mu")
WISHART_KEYS = ("invU_0", "nu_0", "logdet_invU_0", "invU", "U", "nu", "logdet_invU")
MNW_KEYS = ("mu_0", "mu", "invV_0", "invV", "V", "logdetinvV", "logdetinvV_0")


def flatten_state(s, prefix=""):
    """dict-of-dicts of tensors -> flat {name: contiguous tensor} (for np.savez)."""
    out = {}
    for k, v in s.items():
        if isinstance(v, dict):
            out.update(flatten_state(v, prefix + k + "."))
        elif isinstance(v, torch.Tensor):
            out[prefix + k] = v.detach().clone().contiguous()
    return out


def load_state(s, flat, prefix="", dtype=None):
    """Overwrite tensors of ``s`` in place from a flat mapping (inverse of flatten_state)."""
    for k, v in list(s.items()):
        if isinstance(v, dict):
            load_state(v, flat, prefix + k + ".", dtype)
        elif prefix + k in flat:
            t = torch.as_tensor(flat[prefix + k])
            s[k] = t.to(dtype) if (dtype is not None and t.is_floating_point()) else t.clone()
    return s


def to_dtype(s, dtype):
    """Deep-cast every floating tensor of a state dict (fp64 ground-truth runs)."""
    for k, v in list(s.items()):
        if isinstance(v, dict):
            to_dtype(v, dtype)
        elif isinstance(v, torch.Tensor) and v.is_floating_point():
            s[k] = v.to(dtype)
    return s
