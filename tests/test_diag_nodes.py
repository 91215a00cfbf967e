"""Diagonal-precision nodes: NormalGamma / GaussianMixtureModel(isotropic=True) and MatrixNormalGamma /
MixtureofLinearTransforms(type='Gamma').

CPU part: the oracle's restatement of dists/Gamma.py, dists/NormalGamma.py, dists/DiagonalWishart.py and
transforms/MatrixNormalGamma.py is pinned to fixtures produced by the UNMODIFIED reference (tests/golden/make_golden.py).
GPU part (-m gpu): the CUDA path — the streaming diagonal E-step (vbmp_diag_estep), the Gram kernel's diagonal mode,
vbmp_mnw_prep_ex — against the same fixtures and, at a BASELINE-sized shape, against the fp64 oracle.
"""
import numpy as np
import pytest
import torch

import pyvbmp_b200 as V
from oracle import vbem_oracle as O
from _util import load_golden, tag, assert_close, assert_maxabs, argmax_mismatch_report

TIGHT = 2e-5
PARITY = 1e-4
DEV = "cuda:0"
NG_STATE = ("dist.mu", "dist.lambda_mu", "dist.gamma.alpha", "dist.gamma.beta", "pi.alpha")
MNG_STATE = ("mu", "invV", "V", "invU.gamma.alpha", "invU.gamma.beta")


def _flat(o):
    return O.flatten_state(o)


# ---------------------------------------------------------------------------------------------------------------------
# CPU: oracle vs the reference's fixtures
# ---------------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", ["gmm_iso_d8_k6", "gmm_iso_d64_k32"])
def test_oracle_isotropic_gmm_trajectory(name):
    fix = load_golden(name)
    X = torch.as_tensor(fix["X"])
    nc, iters = int(fix["nc"]), int(fix["iters"])
    torch.manual_seed(0)
    m = O.gmm_new(nc, X.shape[-1], isotropic=True)
    O.load_state(m, tag(fix, "init"))
    assert_close(O.ng_elog_like(m["dist"], X.unsqueeze(-2)), fix["init/Elog_like"], TIGHT, "Elog_like init")
    assert_close(O.mixture_kl(m), fix["init/KL"], TIGHT, "KL init")
    trace = O.mixture_update(m, X, 1)
    it1 = tag(fix, "iter1")
    for k in NG_STATE:
        assert_close(_flat(m)[k], it1[k], TIGHT, k + " iter1")
    assert_close(m["NA"], it1["NA"], TIGHT, "NA")
    assert_close(m["logZ"], it1["logZ"], TIGHT, "logZ")
    assert_maxabs(m["p"], it1["p"], 1e-5, "p iter1")
    assert_close(O.mixture_kl(m), it1["KL"], TIGHT, "KL iter1")
    trace += O.mixture_update(m, X, iters - 1)
    got = np.array([float(e) for e in trace])
    assert np.max(np.abs(got - fix["ELBO"]) / np.abs(fix["ELBO"])) < PARITY
    assert (m["p"].argmax(-1).numpy() == fix["final/assignment"]).mean() > 0.999
    O.load_state(m, tag(fix, "final"))
    assert_close(O.mixture_elog_like(m, X), fix["final/Elog_like"], TIGHT, "Elog_like final")
    assert_close(O.mixture_kl(m), fix["final/KL"], TIGHT, "KL final")


def test_oracle_normal_gamma_beta_lr_steps():
    fix = load_golden("ng_beta_lr")
    torch.manual_seed(0)
    s = O.ng_new((3,), (4,), scale=0.7)
    O.load_state(s, tag(fix, "init"))
    for i in range(3):
        X, p = torch.as_tensor(fix[f"X{i}"]), torch.as_tensor(fix[f"p{i}"])
        O.ng_ss_update(s, *O.ng_raw_stats(s, X, p), lr=0.6, beta=0.9)
        ref = tag(fix, f"step{i}")
        for k in ("mu", "lambda_mu", "gamma.alpha", "gamma.beta"):
            assert_close(_flat(s)[k], ref[k], TIGHT, f"{k} step{i}")
        for k in ("SExx", "SEx", "N"):
            assert_close(s[k], ref[k], TIGHT, f"{k} step{i}")
    O.ng_ss_update(s, *O.ng_raw_stats(s, torch.as_tensor(fix["Xb"]), None), lr=1.0, beta=None)
    ref = tag(fix, "pnone")
    for k in ("mu", "lambda_mu", "gamma.alpha", "gamma.beta"):
        assert_close(_flat(s)[k], ref[k], TIGHT, f"{k} p=None")
    assert_close(O.ng_kl(s), fix["final/KL"], TIGHT, "KL")
    assert_close(O.ng_elog_like(s, torch.as_tensor(fix["X2"])), fix["final/Elog_like"], TIGHT, "Elog_like")


@pytest.mark.parametrize("pad", [1, 0])
def test_oracle_matrix_normal_gamma_steps(pad):
    fix = load_golden(f"mng_n4_p5_k3_pad{pad}")
    n, p, K = int(fix["n"]), int(fix["p"]), int(fix["K"])
    torch.manual_seed(0)
    s = O.mng_new((n, p), (K,), scale=0.8, pad_X=bool(pad))
    O.load_state(s, tag(fix, "init"))
    X, Y, r = (torch.as_tensor(fix[k]) for k in ("X", "Y", "r"))
    assert_close(O.mng_elog_like(s, X, Y), fix["init/Elog_like"], TIGHT, "Elog_like init")
    assert_close(O.mng_kl(s), fix["init/KL"], TIGHT, "KL init")
    O.mng_ss_update(s, *O.mnw_raw_stats_exact(s, X, Y, r), lr=1.0, beta=None)
    ref = tag(fix, "step0")
    for k in MNG_STATE:
        assert_close(_flat(s)[k], ref[k], 5e-5, k + " step0")
    assert_close(O.mng_elog_like(s, X, Y), fix["step0/Elog_like"], 5e-5, "Elog_like step0")
    assert_close(O.mng_kl(s), fix["step0/KL"], 5e-5, "KL step0")
    O.mng_ss_update(s, *O.mnw_raw_stats_exact(s, X, Y, r), lr=0.5, beta=0.8)
    ref = tag(fix, "step1")
    for k in MNG_STATE:
        assert_close(_flat(s)[k], ref[k], 5e-5, k + " step1")
    assert_close(O.mng_kl(s), fix["step1/KL"], 5e-5, "KL step1")


@pytest.mark.parametrize("name", ["molt_gamma_n3_p4_k5", "molt_gamma_n32_p32_k8"])
def test_oracle_molt_gamma_trajectory(name):
    fix = load_golden(name)
    n, p, K, iters = (int(fix[k]) for k in ("n", "p", "K", "iters"))
    torch.manual_seed(0)
    m = O.molt_new(n, p, K, type='Gamma')
    O.load_state(m, tag(fix, "init"))
    X, Y = torch.as_tensor(fix["X"]).unsqueeze(-1), torch.as_tensor(fix["Y"]).unsqueeze(-1)
    trace = O.molt_raw_update(m, X, Y, 1)
    it1 = tag(fix, "iter1")
    assert_maxabs(m["p"], it1["p"], 2e-5, "p iter1")
    for k in ("W.mu", "W.invV", "W.V", "W.invU.gamma.alpha", "W.invU.gamma.beta", "pi.alpha"):
        assert_close(_flat(m)[k], it1[k], 5e-5, k)
    trace += O.molt_raw_update(m, X, Y, iters - 1)
    got = np.array([float(e) for e in trace])
    assert np.max(np.abs(got - fix["ELBO"]) / np.abs(fix["ELBO"])) < PARITY
    assert (m["p"].argmax(-1).numpy() == fix["final/assignment"]).mean() > 0.995


def test_diag_node_constructors_match_the_oracle_rng_order():
    """Same seed -> same initial state (the mirrors consume the global RNG in the reference's order)."""
    torch.manual_seed(5)
    a = V.GaussianMixtureModel(4, 3, isotropic=True)
    torch.manual_seed(5)
    b = O.gmm_new(4, 3, isotropic=True)
    assert torch.equal(a.dist.mu, b["dist"]["mu"]) and torch.equal(a.dist.gamma.beta, b["dist"]["gamma"]["beta"])
    assert torch.equal(a.pi.alpha, b["pi"]["alpha"]) and torch.equal(a.dist.lambda_mu, b["dist"]["lambda_mu"])
    torch.manual_seed(6)
    c = V.MixtureofLinearTransforms(3, 4, 5, type='Gamma')
    torch.manual_seed(6)
    d = O.molt_new(3, 4, 5, type='Gamma')
    assert torch.equal(c.W.mu, d["W"]["mu"]) and torch.equal(c.W.invU.gamma.alpha, d["W"]["invU"]["gamma"]["alpha"])
    assert torch.equal(c.pi.alpha, d["pi"]["alpha"])
    with pytest.raises(ValueError):
        V.MixtureofLinearTransforms(3, 4, 5, type='Cauchy')


# ---------------------------------------------------------------------------------------------------------------------
# GPU: the CUDA path vs the fixtures and the fp64 oracle
# ---------------------------------------------------------------------------------------------------------------------

def _set(obj, flat, device=DEV):
    for k, v in flat.items():
        parts = k.split(".")
        o = obj
        ok = True
        for a in parts[:-1]:
            if not hasattr(o, a):
                ok = False
                break
            o = getattr(o, a)
        if ok and isinstance(getattr(o, parts[-1], None), (torch.Tensor, float)) and isinstance(v, torch.Tensor):
            setattr(o, parts[-1], v.to(device))


def _get(obj, path):
    for a in path.split("."):
        obj = getattr(obj, a)
    return obj


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["gmm_iso_d8_k6", "gmm_iso_d64_k32"])
def test_isotropic_gmm_golden(name):
    from pyvbmp_b200 import _lib
    fix = load_golden(name)
    X = torch.as_tensor(fix["X"]).to(DEV)
    nc, iters = int(fix["nc"]), int(fix["iters"])
    torch.manual_seed(0)
    m = V.GaussianMixtureModel(nc, X.shape[-1], isotropic=True).to(DEV)
    _set(m, tag(fix, "init"))
    n0 = _lib.LAUNCHES
    ll = m.dist.Elog_like(X.unsqueeze(-2))
    assert _lib.LAUNCHES > n0
    assert_close(ll, fix["init/Elog_like"], 2e-6, "Elog_like init")
    assert_close(m.KLqprior(), fix["init/KL"], PARITY, "KL init")
    m.update(X, 1)
    it1 = tag(fix, "iter1")
    assert abs(float(m.ELBO_last) - fix["ELBO"][0]) <= PARITY * abs(fix["ELBO"][0])
    for k in NG_STATE:
        assert_close(_get(m, k), it1[k], PARITY, k + " iter1")
    assert_close(m.NA, it1["NA"], PARITY, "NA")
    assert_close(m.logZ, it1["logZ"], PARITY, "logZ")
    assert_maxabs(m.p.cpu(), it1["p"], 1e-4, "p iter1")
    nbad, margins = argmax_mismatch_report(m.p, it1["p"])
    assert nbad == 0, (nbad, margins)
    elbo = [float(m.ELBO_last)]
    for _ in range(iters - 1):
        m.update(X, 1)
        elbo.append(float(m.ELBO_last))
    assert np.max(np.abs(np.array(elbo) - fix["ELBO"]) / np.abs(fix["ELBO"])) < PARITY
    assert (m.assignment().cpu().numpy() == fix["final/assignment"]).mean() > 0.999
    _set(m, tag(fix, "final"))
    assert_close(m.Elog_like(X), fix["final/Elog_like"], 2e-6, "Elog_like final")
    assert_close(m.KLqprior(), fix["final/KL"], PARITY, "KL final")


@pytest.mark.gpu
def test_normal_gamma_beta_lr_steps_gpu():
    fix = load_golden("ng_beta_lr")
    torch.manual_seed(0)
    s = V.NormalGamma((3,), (4,), scale=0.7).to(DEV)
    _set(s, tag(fix, "init"))
    for i in range(3):
        s.raw_update(torch.as_tensor(fix[f"X{i}"]).to(DEV), torch.as_tensor(fix[f"p{i}"]).to(DEV), lr=0.6, beta=0.9)
        ref = tag(fix, f"step{i}")
        for k in ("mu", "lambda_mu", "gamma.alpha", "gamma.beta", "SExx", "SEx", "N"):
            assert_close(_get(s, k), ref[k], PARITY, f"{k} step{i}")
    s.raw_update(torch.as_tensor(fix["Xb"]).to(DEV), None, lr=1.0, beta=None)
    ref = tag(fix, "pnone")
    for k in ("mu", "lambda_mu", "gamma.alpha", "gamma.beta"):
        assert_close(_get(s, k), ref[k], PARITY, f"{k} p=None")
    assert_close(s.KLqprior(), fix["final/KL"], PARITY, "KL")
    assert_close(s.Elog_like(torch.as_tensor(fix["X2"]).to(DEV)), fix["final/Elog_like"], PARITY, "Elog_like")


@pytest.mark.gpu
@pytest.mark.parametrize("pad", [1, 0])
def test_matrix_normal_gamma_steps_gpu(pad):
    fix = load_golden(f"mng_n4_p5_k3_pad{pad}")
    n, p, K = int(fix["n"]), int(fix["p"]), int(fix["K"])
    torch.manual_seed(0)
    s = V.MatrixNormalGamma((n, p), (K,), scale=0.8, pad_X=bool(pad)).to(DEV)
    _set(s, tag(fix, "init"))
    X, Y, r = (torch.as_tensor(fix[k]).to(DEV) for k in ("X", "Y", "r"))
    assert_close(s.Elog_like(X, Y), fix["init/Elog_like"], PARITY, "Elog_like init")
    assert_close(s.KLqprior(), fix["init/KL"], PARITY, "KL init")
    s.raw_update(X, Y, p=r, lr=1.0, beta=None)
    ref = tag(fix, "step0")
    for k in MNG_STATE:
        assert_close(_get(s, k), ref[k], PARITY, k + " step0")
    assert_maxabs(s.logdetinvV.cpu(), ref["logdetinvV"], 1e-4, "logdetinvV")
    assert_close(s.Elog_like(X, Y), fix["step0/Elog_like"], PARITY, "Elog_like step0")
    assert_close(s.KLqprior(), fix["step0/KL"], PARITY, "KL step0")
    s.raw_update(X, Y, p=r, lr=0.5, beta=0.8)
    ref = tag(fix, "step1")
    for k in MNG_STATE + ("SExx", "SEyx", "SEyy", "N"):
        assert_close(_get(s, k), ref[k], PARITY, k + " step1")
    assert_close(s.KLqprior(), fix["step1/KL"], PARITY, "KL step1")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["molt_gamma_n3_p4_k5", "molt_gamma_n32_p32_k8"])
def test_molt_gamma_golden(name):
    fix = load_golden(name)
    n, p, K, iters = (int(fix[k]) for k in ("n", "p", "K", "iters"))
    torch.manual_seed(0)
    m = V.MixtureofLinearTransforms(n, p, K, type='Gamma').to(DEV)
    _set(m, tag(fix, "init"))
    X, Y = torch.as_tensor(fix["X"]).unsqueeze(-1).to(DEV), torch.as_tensor(fix["Y"]).unsqueeze(-1).to(DEV)
    m.raw_update(X, Y, iters=1)
    it1 = tag(fix, "iter1")
    assert abs(float(m.ELBO_last) - fix["ELBO"][0]) <= PARITY * abs(fix["ELBO"][0])
    assert_maxabs(m.p.cpu(), it1["p"], 5e-4, "p iter1")
    assert_close(m.logZ, it1["logZ"], PARITY, "logZ_n")
    for k in ("W.mu", "W.invV", "W.V", "W.invU.gamma.alpha", "W.invU.gamma.beta", "pi.alpha"):
        assert_close(_get(m, k), it1[k], PARITY, k)
    elbo = [float(m.ELBO_last)]
    for _ in range(iters - 1):
        m.raw_update(X, Y, iters=1)
        elbo.append(float(m.ELBO_last))
    assert np.max(np.abs(np.array(elbo) - fix["ELBO"]) / np.abs(fix["ELBO"])) < PARITY
    assert (m.assignment().cpu().numpy() == fix["final/assignment"]).mean() > 0.995
    pY, pr = m.predict(X)                                         # generic (reference op order) predict on the Gamma node
    assert pr.shape == (X.shape[0], K) and bool(torch.isfinite(pY.mean()).all())


@pytest.mark.gpu
def test_isotropic_gmm_cfg2_shape_vs_fp64_oracle():
    """d = 64, K = 256, N = 65 536 with NormalGamma components: streaming E-step + the Gram kernel's diagonal mode on tcgen05
    (with the K2 -> K3 hand-over absent: the diagonal E-step does not write operand images) against the fp64 oracle."""
    from pyvbmp_b200 import _lib
    N, K, d = 65536, 256, 64
    g = torch.Generator().manual_seed(77)
    mu = 1.5 * torch.randn(K, d, generator=g)
    sd = 0.5 + torch.rand(K, d, generator=g)
    z = torch.randint(K, (N,), generator=g)
    X = mu[z] + sd[z] * torch.randn(N, d, generator=g)
    torch.manual_seed(7)
    m = V.GaussianMixtureModel(K, d, isotropic=True)
    m.initialize(X)
    ref = O.gmm_new(K, d, isotropic=True)
    O.load_state(ref, {"dist.mu": m.dist.mu.clone(), "dist.lambda_mu": m.dist.lambda_mu.clone(),
                       "dist.gamma.alpha": m.dist.gamma.alpha.clone(), "dist.gamma.beta": m.dist.gamma.beta.clone(),
                       "pi.alpha": m.pi.alpha.clone()})
    O.to_dtype(ref, torch.float64)
    m.to(DEV)
    Xd, X64 = X.to(DEV), X.double()
    for it in range(3):
        _set(m, {k: v.float() for k, v in O.flatten_state(ref).items()})
        m.update(Xd, 1)
        tr = O.mixture_update(ref, X64, 1, chunk=4096)
        assert abs(float(m.ELBO_last) - float(tr[0])) <= PARITY * abs(float(tr[0])), it
        L = float(ref["log_p"].max(-1)[0].abs().max())
        assert_maxabs(m.p.cpu().double(), ref["p"], max(2e-5, 4e-7 * L), f"p it{it} (|logit| {L:.1e})")
        nbad, margins = argmax_mismatch_report(m.p, ref["p"], ref["log_p"])
        assert nbad == 0 or max(margins) < 1e-3, (it, nbad, margins)
        assert_close(m.NA, ref["NA"], PARITY, "NA")
        flat = O.flatten_state(ref)
        for k in NG_STATE:
            assert_close(_get(m, k), flat[k], PARITY, f"{k} it{it}")
    # the diagonal statistics from the tensor-core kernel equal an fp64 evaluation per component
    pw = m.p.double()
    SEx = pw.t() @ Xd.double()
    SExx = pw.t() @ (Xd.double() ** 2)
    a, b, c = m.dist._stats(Xd.view(N, 1, d), m.p)
    assert_close(b, SEx, 1e-5, "SEx")
    assert float(((a.double() - SExx).abs() / SExx.abs().clamp_min(1e-30)).max()) < 2e-5
    assert_close(c, pw.sum(0), 1e-5, "N")
    # the diagonal E-step handed the responsibilities over pre-split (K2 -> K3): same statistics, bit for bit, as from a copy
    # of p, which the Gram call has to split itself
    m.update_assignments(Xd)
    assert _lib._rpack_key(Xd.device) in _lib._rpack_rec
    s1 = [t.clone() for t in m.dist._stats(Xd.view(N, 1, d), m.p)]
    s2 = m.dist._stats(Xd.view(N, 1, d), m.p.clone())
    for u, v in zip(s1, s2):
        assert torch.equal(u, v)


@pytest.mark.gpu
@pytest.mark.parametrize("N,d,K", [(5000, 64, 32), (4100, 32, 64), (3000, 96, 16), (6000, 64, 132)])
def test_diagonal_statistics_kernel(N, d, K):
    """The Gram kernels' diagonal mode (vbmp_gram flags bit 1) on both tensor-core variants — components on the MMA's N
    dimension for K <= 64, on M above — against an fp64 evaluation, per component."""
    from pyvbmp_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(N + K)
    X = (torch.randn(N, d, generator=g, device=DEV) * 1.7 + 0.3).contiguous()
    P = torch.softmax(2.0 * torch.randn(N, K, generator=g, device=DEV), -1).contiguous()
    xg = torch.zeros(1, dtype=torch.int32, device=DEV)
    G = _lib.gram(X.view(N, 1, d), None, N, 1, xg, P.view(N, 1, K), 1, xg, 1, K, _lib.pad_dim(d), diag=True).view(K, d + 1, d + 1)
    Pd, Xd = P.double(), X.double()
    SExx, SEx, Nk = Pd.t() @ Xd ** 2, Pd.t() @ Xd, Pd.sum(0)
    assert float(((G[:, :d, :d].diagonal(dim1=-2, dim2=-1).double() - SExx).abs() / SExx).max()) < 1e-5
    assert float((G[:, :d, d].double() - SEx).abs().max() / SEx.abs().max()) < 1e-5
    assert float(((G[:, d, d].double() - Nk).abs() / Nk).max()) < 1e-5
    assert torch.equal(G[:, :d, d], G[:, d, :d])


@pytest.mark.gpu
@pytest.mark.parametrize("N,d,K,G", [(777, 5, 6, 1), (1300, 19, 3, 4), (4099, 64, 260, 1), (130, 128, 17, 1)])
def test_diag_estep_kernel_general_shapes(N, d, K, G):
    """vbmp_diag_estep on ragged shapes (N not a multiple of the 128-row tile, K not a multiple of 4 / 128, replica groups
    with their own data columns, d = 128 with a single resident parameter tile) against an fp64 evaluation."""
    from pyvbmp_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(N + d)
    X = (1.5 * torch.randn(N, G, d, generator=g, device=DEV) + 0.2).contiguous()
    mu = torch.randn(G * K, d, generator=g, device=DEV)
    tau = (0.2 + torch.rand(G * K, d, generator=g, device=DEV)).contiguous()
    cst = torch.randn(G * K, generator=g, device=DEV)
    xg = torch.arange(G, dtype=torch.int32, device=DEV)
    L = cst.double().view(1, G, K) - 0.5 * ((X.double().unsqueeze(2) - mu.double().view(1, G, K, d)) ** 2
                                            * tau.double().view(1, G, K, d)).sum(-1)
    lg = _lib.diag_estep(X, N, G, xg, mu, tau, cst, G, K, d, 0)
    scale = float(L.abs().max())
    assert float((lg.double() - L).abs().max()) <= 2e-6 * scale
    p, lzn, NA, lZ = _lib.diag_estep(X, N, G, xg, mu, tau, cst, G, K, d, 1)
    lz = torch.logsumexp(L, -1)
    P = (L - lz.unsqueeze(-1)).exp()
    assert float((lzn.double() - lz).abs().max()) <= 2e-6 * scale
    assert float((p.double() - P).abs().max()) <= 1e-4
    assert float(((NA.double() - P.sum(0)).abs() / P.sum(0).clamp_min(1.0)).max()) <= 1e-4
    assert float(((lZ.double() - lz.sum(0)).abs() / lz.sum(0).abs().clamp_min(1.0)).max()) <= 1e-5


# (No replica-batch Mixture test for NormalGamma: dists/NormalGamma.py:88-94 adds `self.gamma.KLqprior().sum(-1)` — the Gamma
# KL summed over the LAST BATCH dim as well — to a per-component vector, which broadcasts only for batch_shape = (K,); with
# batch (G, K) the reference itself raises.  The mirror and the oracle reproduce that expression, quirk included.)


@pytest.mark.gpu
@pytest.mark.parametrize("N,d,K", [(500, 1, 1), (4000, 1, 3), (700, 3, 2), (3000, 64, 1), (2600, 64, 3), (3000, 128, 2), (9, 2, 1)])
def test_isotropic_gmm_degenerate_shapes(N, d, K):
    """GaussianMixtureModel(isotropic=True) with one feature and with one / two / three components (diagonal E-step, the Gram
    kernel's diagonal mode at K below its 4-component granule, D = 128): three step-wise iterations against the fp64 oracle."""
    g = torch.Generator().manual_seed(5 * N + d + K)
    mu = 2.0 * torch.randn(K, d, generator=g)
    X = mu[torch.randint(K, (N,), generator=g)] + (0.5 + torch.rand(K, d, generator=g))[torch.randint(K, (N,), generator=g)] * torch.randn(N, d, generator=g)
    torch.manual_seed(7)
    m = V.GaussianMixtureModel(K, d, isotropic=True)
    m.initialize(X)
    ref = O.gmm_new(K, d, isotropic=True)
    O.load_state(ref, {"dist.mu": m.dist.mu.clone(), "dist.lambda_mu": m.dist.lambda_mu.clone(),
                       "dist.gamma.alpha": m.dist.gamma.alpha.clone(), "dist.gamma.beta": m.dist.gamma.beta.clone(),
                       "pi.alpha": m.pi.alpha.clone()})
    O.to_dtype(ref, torch.float64)
    m.to(DEV)
    Xd, X64 = X.to(DEV), X.double()
    for it in range(3):
        _set(m, {k: v.float() for k, v in O.flatten_state(ref).items()})
        m.update(Xd, 1)
        tr = O.mixture_update(ref, X64, 1)
        assert m.p.shape == (N, K)
        assert abs(float(m.ELBO_last) - float(tr[0])) <= PARITY * abs(float(tr[0])), it
        L = float(ref["log_p"].abs().max())
        assert_maxabs(m.p.cpu().double(), ref["p"], max(2e-5, 4e-7 * L), f"p it{it} (|logit| {L:.1e})")
        assert_close(m.NA, ref["NA"], PARITY, "NA")
        flat = O.flatten_state(ref)
        for k in NG_STATE:
            assert_close(_get(m, k), flat[k], PARITY, f"{k} it{it}")
