"""Model-level GPU parity at the BASELINE.json shapes (SURVEY.md Appendix D rows that the small golden fixtures do
not reach): MixtureofLinearTransforms n = p = 32, K = 64, N = 65 536 (cfg3 shape) and ARHMM K = 32, n = p = 16,
T = 64, S = 64 (cfg4 shape), step-wise against the fp64 oracle.  These are the cases where mnw_prep, the tcgen05
E-step, the tcgen05 Gram (D = 64 / 32 with a two-part z = [x; y]) and mnw_update meet the oracle TOGETHER; plus the
standalone Wishart entry points.
"""
import pytest
import torch

import pyvbmp_b200 as V
from oracle import vbem_oracle as O
from _util import assert_close, argmax_mismatch_report, assert_maxabs
from test_cuda_parity import set_state, get

pytestmark = pytest.mark.gpu
PARITY = 1e-4
DEV = "cuda:0"

MOLT_STATE = ("W.mu", "W.invV", "W.V", "W.invU.invU", "W.invU.U", "W.invU.nu", "pi.alpha")


def _cfg3_data(N, n=32, p=32, K=64, seed=11):
    """SURVEY.md §8d cfg3 recipe: X ~ N(0, I), W_k = randn / sqrt(p), b_k = randn, Y = W_z x + b_z + 0.1 randn."""
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(N, p, generator=g)
    W = torch.randn(K, n, p, generator=g) / p ** 0.5
    b = torch.randn(K, n, generator=g)
    z = torch.randint(K, (N,), generator=g)
    Y = torch.einsum("nij,nj->ni", W[z], X) + b[z] + 0.1 * torch.randn(N, n, generator=g)
    return X.unsqueeze(-1), Y.unsqueeze(-1)


def _p_gate(p_gpu, ref, what):
    """Responsibilities against the fp64 oracle: 2e-5 where the logits are O(1e2) or smaller (converged states; the
    reference's own fp32 noise there is 1.6e-5), otherwise the fp32 floor eps32 * |logit|."""
    L = float(ref["log_p"].max(-1)[0].abs().max())
    err = float((p_gpu.cpu().double() - ref["p"]).abs().max())
    print(f"[p-gate] {what}: max |dp| = {err:.2e} at |logit| <= {L:.2e}")
    assert_maxabs(p_gpu.cpu().double(), ref["p"], max(2e-5, 2e-7 * L), f"{what} (|logit| {L:.2e})")
    nbad, margins = argmax_mismatch_report(p_gpu, ref["p"], ref["log_p"])
    assert nbad == 0 or max(margins) < 1e-3, (what, nbad, margins)
    return L


def test_molt_cfg3_shape_vs_fp64_oracle():
    """MoLT n = p = 32, K = 64, N = 65 536: three step-wise E+M iterations (state copied from the oracle before each)."""
    N, n, p, K = 65536, 32, 32, 64
    X, Y = _cfg3_data(N, n, p, K)
    torch.manual_seed(5)
    m = V.MixtureofLinearTransforms(n, p, K)
    ref = O.molt_new(n, p, K)
    O.load_state(ref, {"W.mu": m.W.mu.clone(), "pi.alpha": m.pi.alpha.clone()})
    O.to_dtype(ref, torch.float64)
    m.to(DEV)
    Xd, Yd, X64, Y64 = X.to(DEV), Y.to(DEV), X.double(), Y.double()
    from pyvbmp_b200 import _lib
    for it in range(3):
        set_state(m, {k: v.float() for k, v in O.flatten_state(ref).items()})
        n0 = _lib.LAUNCHES
        m.raw_update(Xd, Yd, iters=1)
        assert _lib.LAUNCHES > n0
        tr = O.molt_raw_update(ref, X64, Y64, 1, exact=False, chunk=8192)
        assert abs(float(m.ELBO_last) - float(tr[0])) <= PARITY * abs(float(tr[0])), it
        _p_gate(m.p, ref, f"p it{it}")
        assert_close(m.logZ, ref["logZ"], PARITY, f"logZ_n it{it}")
        flat = O.flatten_state(ref)
        for k in MOLT_STATE:
            # iteration 0 starts from O(1e3) logits under the broad prior: the reference's own fp32 run is ~1.5e-4 from its
            # fp64 run there; from iteration 1 on the 1e-4 gate applies against the fp64 truth
            assert_close(get(m, k), flat[k], 3e-4 if it == 0 else PARITY, f"{k} it{it}")
        assert_maxabs(m.W.logdetinvV.cpu().double(), flat["W.logdetinvV"], 2e-3, "logdetinvV")
        assert int((m.W.info != 0).sum()) == 0
    # converged-state E-step: a few free-running iterations of the oracle, then one E-step on its state
    O.molt_raw_update(ref, X64, Y64, 3, exact=False, chunk=8192)
    set_state(m, {k: v.float() for k, v in O.flatten_state(ref).items()})
    O.molt_update_assignments(ref, X64, Y64, exact=False, chunk=8192)
    m.update_assignments(Xd, Yd)
    L = _p_gate(m.p, ref, "p converged")
    assert L < 2e2, L                                     # i.e. the 2e-5 gate was the one applied
    assert_close(m.logZ, ref["logZ"], PARITY, "logZ_n converged")
    assert (m.assignment().cpu() == ref["p"].argmax(-1)).float().mean() > 0.9999


def test_arhmm_cfg4_shape_vs_fp64_oracle():
    """ARHMM K = 32, n = p = 16, T = 64, S = 64 (tcgen05 E-step and Gram on z = [x; y], D = 32; forward-backward kernel)."""
    K, n, T, S = 32, 16, 64, 64
    g = torch.Generator().manual_seed(3)
    # a switching AR(1): A_k = 0.95 Q_k (random orthogonal), sticky transitions (tests/test_models.py:20-28)
    Q = torch.linalg.qr(torch.randn(K, n, n, generator=g))[0] * 0.95
    trans = 4.0 * torch.eye(K) + torch.rand(K, K, generator=g)
    trans = trans / trans.sum(-1, True)
    y = torch.zeros(T + 1, S, n)
    y[0] = torch.randn(S, n, generator=g)
    z = torch.randint(K, (S,), generator=g)
    for t in range(T):
        z = torch.multinomial(trans[z], 1, generator=g).squeeze(-1)
        y[t + 1] = torch.einsum("sij,sj->si", Q[z], y[t]) + 0.3 * torch.randn(S, n, generator=g)
    X = y[:-1].reshape(T, S, 1, n, 1).contiguous()
    Y = y[1:].reshape(T, S, 1, n, 1).contiguous()
    torch.manual_seed(9)
    h = V.ARHMM(K, n, n)
    ref = O.arhmm_new(K, n, n)
    O.load_state(ref, {"obs.mu": h.obs_dist.mu.clone(), "transition.alpha": h.transition.alpha.clone(),
                       "initial.alpha": h.initial.alpha.clone()})
    O.to_dtype(ref, torch.float64)
    h.to(DEV)
    Xd, Yd, X64, Y64 = X.to(DEV), Y.to(DEV), X.double(), Y.double()
    keys = ("obs.mu", "obs.invV", "obs.V", "obs.invU.invU", "obs.invU.U", "obs.invU.nu", "transition.alpha", "initial.alpha")
    for it in range(3):
        set_state(h, {k.replace("obs.", "obs_dist."): v.float() for k, v in O.flatten_state(ref).items()})
        ol = h.obs_logits((Xd, Yd))
        ol_ref = O.arhmm_obs_logits(ref, X64, Y64, exact=False)
        assert ol.shape == ol_ref.shape
        # logits relative to their own magnitude (fp32 floor), and absolutely near each row's maximum
        assert_close(ol, ol_ref, 2e-6, f"obs_logits it{it}")
        h.update((Xd, Yd), iters=1)
        tr = O.arhmm_update(ref, X64, Y64, 1, exact=False)
        L = float(ol_ref.abs().max())
        assert_maxabs(h.p.cpu().double(), ref["p"], max(5e-5, 1e-6 * L), f"p after forward-backward it{it} (|logit| {L:.1e})")
        assert_close(h.logZ, ref["logZ"], PARITY, f"logZ it{it}")
        assert_close(h.NA, ref["NA"], PARITY, f"NA it{it}")
        assert abs(float(h.ELBO_last) - float(tr[0])) <= PARITY * abs(float(tr[0])), it
        flat = O.flatten_state(ref)
        for k in keys:
            assert_close(get(h, k.replace("obs.", "obs_dist.")), flat[k], 3e-4 if it == 0 else PARITY, f"{k} it{it}")


def test_gmm_d128_vs_fp64_oracle():
    """GaussianMixtureModel d = 128 (the north star's upper feature dimension), K = 64, N = 16 384: the tcgen05 kernels at
    Dp = 128 — E-step with two accumulator buffers, Gram over 8385 pair columns — step-wise against the fp64 oracle."""
    from pyvbmp_b200 import _lib
    N, K, d = 16384, 64, 128
    g = torch.Generator().manual_seed(13)
    mu = 0.4 * torch.randn(K, d, generator=g)
    A = torch.eye(d) + 0.3 * torch.randn(K, d, d, generator=g) / 11
    z = torch.randint(K, (N,), generator=g)
    X = mu[z] + torch.einsum("nij,nj->ni", A[z], torch.randn(N, d, generator=g))
    torch.manual_seed(7)
    m = V.GaussianMixtureModel(K, d)
    m.initialize(X)
    ref = O.gmm_new(K, d)
    O.load_state(ref, {"dist.mu": m.dist.mu.clone(), "pi.alpha": m.pi.alpha.clone()})
    O.to_dtype(ref, torch.float64)
    m.to(DEV)
    Xd, X64 = X.to(DEV), X.double()
    keys = ("dist.mu", "dist.lambda_mu", "dist.invU.invU", "dist.invU.U", "dist.invU.nu", "pi.alpha")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error")                   # leaving the tensor-core window would warn: it must not
        for it in range(3):
            set_state(m, {k: v.float() for k, v in O.flatten_state(ref).items()})
            m.update(Xd, 1)
            tr = O.mixture_update(ref, X64, 1, exact=False, chunk=2048)
            assert abs(float(m.ELBO_last) - float(tr[0])) <= PARITY * abs(float(tr[0])), it
            L = float(ref["log_p"].max(-1)[0].abs().max())
            assert_maxabs(m.p.cpu().double(), ref["p"], max(2e-4, 4e-7 * L), f"p it{it} (|logit| {L:.2e})")
            nbad, margins = argmax_mismatch_report(m.p, ref["p"], ref["log_p"])
            assert nbad == 0 or max(margins) < 1e-3, (it, nbad, margins)
            assert_close(m.NA, ref["NA"], PARITY, "NA")
            flat = O.flatten_state(ref)
            for k in keys:
                assert_close(get(m, k), flat[k], 3e-4 if (it == 0 or k == "dist.invU.U") else PARITY, f"{k} it{it}")
            m.dist.invU.check()


def test_wishart_standalone_update_and_kl():
    """Wishart.ss_update / KLqprior / ElogdetinvSigma called directly (dists/Wishart.py:43-56, 82-94): SURVEY.md §8 a9."""
    d, K = 24, 7
    g = torch.Generator().manual_seed(2)
    torch.manual_seed(0)
    w = V.Wishart((d, d), (K,), scale=torch.tensor(0.6)).to(DEV)
    ref = O.wishart_new(d, (K,), scale=0.6, dtype=torch.float64)
    for step, (lr, beta) in enumerate([(1.0, None), (0.5, None), (0.7, 0.9)]):
        A = torch.randn(K, 40, d, generator=g)
        SExx = A.transpose(-1, -2) @ A
        Nk = torch.full((K,), 40.0) + torch.rand(K, generator=g)
        w.ss_update(SExx.to(DEV), Nk.to(DEV), lr=lr, beta=beta)
        O.wishart_ss_update(ref, SExx.double(), Nk.double(), lr=lr, beta=beta)
        for k in ("invU", "U", "nu"):
            assert_close(getattr(w, k), ref[k], PARITY, f"{k} step{step}")
        assert_maxabs(w.logdet_invU.cpu().double(), ref["logdet_invU"], 1e-4 * d, f"logdet step{step}")
        assert_close(w.KLqprior(), O.wishart_kl(ref), PARITY, f"KL step{step}")
        assert_close(w.ElogdetinvSigma(), O.wishart_ElogdetinvSigma(ref), PARITY, f"ElogdetinvSigma step{step}")
        w.check()
    # a tensor-valued (per-component) scale, as the reference's constructor accepts (dists/Wishart.py:9-26)
    sc = (0.5 + torch.rand(K, 1, 1, generator=g))
    w2 = V.Wishart((d, d), (K,), scale=sc)
    assert_close(w2.invU_0, sc ** 2 * torch.eye(d), 1e-7, "invU_0 per-component scale")
    assert_close(w2.logdet_invU_0, (sc ** 2 * torch.eye(d)).logdet(), 1e-5, "logdet_invU_0 per-component scale")
    assert_close(w2.U, (sc ** 2 * torch.eye(d)).inverse(), 1e-6, "U per-component scale")


def test_molt_update_given_beliefs_vs_fp64_oracle():
    """Expectation-input E and M steps at a shape inside the kernels' windows: MoLT n = p = 16, K = 8,
    N = 8192 with per-sample Gaussian beliefs about inputs and outputs.  Elog_like_given_pX_pY runs K1 + K2 on the means
    plus vbmp_rowterm on the flattened covariances; update(pX, pY) runs K3 on the means plus vbmp_wsum (the Gram kernel's
    "lin" mode) on the covariances.  Against the fp64 oracle, two step-wise iterations."""
    N, n, p, K = 8192, 16, 16, 8
    X, Y = _cfg3_data(N, n, p, K, seed=21)
    g = torch.Generator().manual_seed(22)
    Ax = 0.2 * torch.randn(N, p, p, generator=g)
    Ay = 0.2 * torch.randn(N, n, n, generator=g)
    Sx = Ax @ Ax.transpose(-1, -2) + 0.05 * torch.eye(p)
    Sy = Ay @ Ay.transpose(-1, -2) + 0.05 * torch.eye(n)
    torch.manual_seed(6)
    m = V.MixtureofLinearTransforms(n, p, K)
    ref = O.molt_new(n, p, K)
    O.load_state(ref, {"W.mu": m.W.mu.clone(), "pi.alpha": m.pi.alpha.clone()})
    O.to_dtype(ref, torch.float64)
    m.to(DEV)
    pX = V.MultivariateNormal_vector_format(mu=X.to(DEV), Sigma=Sx.to(DEV))
    pY = V.MultivariateNormal_vector_format(mu=Y.to(DEV), Sigma=Sy.to(DEV))
    from pyvbmp_b200 import _lib
    for it in range(2):
        set_state(m, {k: v.float() for k, v in O.flatten_state(ref).items()})
        _lib.profile_begin(64)
        m.update(pX, pY, iters=1)
        prof = _lib.profile_end()
        assert "vbmp_wsum" in prof and "vbmp_rowterm" in prof, sorted(prof)      # the kernels ran, not the torch branch
        elbo = O.molt_update_given(ref, X.double(), Sx.double(), Y.double(), Sy.double())
        assert abs(float(m.ELBO_last) - float(elbo)) <= PARITY * abs(float(elbo)), it
        assert_maxabs(m.p.cpu().double(), ref["p"], 2e-4 if it == 0 else 2e-5, f"p it{it}")
        assert_close(m.logZ, ref["logZ"], PARITY, f"logZ_n it{it}")
        flat = O.flatten_state(ref)
        for k in MOLT_STATE:
            assert_close(get(m, k), flat[k], 3e-4 if it == 0 else PARITY, f"{k} it{it}")


def test_cfg3_full_size_properties():
    """BASELINE.json configs[2] at its full size (MixtureofLinearTransforms, N = 8 388 608, n = p = 32, K = 64): properties that
    need no oracle, plus an fp64 evaluation of one component's statistics from the kernel's own responsibilities."""
    import numpy as np
    N, n, p, K = 8_388_608, 32, 32, 64
    g = torch.Generator(device=DEV).manual_seed(1)
    X = torch.randn(N, p, generator=g, device=DEV)
    W = torch.randn(K, n, p, generator=g, device=DEV) / p ** 0.5
    b = torch.randn(K, n, generator=g, device=DEV)
    z = torch.randint(K, (N,), generator=g, device=DEV)
    Y = torch.empty(N, n, device=DEV)
    for a in range(0, N, 1 << 20):
        e = a + (1 << 20)
        Y[a:e] = torch.einsum("nij,nj->ni", W[z[a:e]], X[a:e]) + b[z[a:e]] + 0.1 * torch.randn(e - a, n, generator=g, device=DEV)
    del W, b, z
    torch.manual_seed(0)
    m = V.MixtureofLinearTransforms(n, p, K, pad_X=True).to(DEV)
    Xc, Yc = X.unsqueeze(-1), Y.unsqueeze(-1)
    elbo = []
    for _ in range(3):
        m.raw_update(Xc, Yc, iters=1, lr=1)
        elbo.append(float(m.ELBO_last))
    assert all(np.isfinite(elbo))
    assert elbo[1] >= elbo[0] - 1e-6 * abs(elbo[0]) and elbo[2] >= elbo[1] - 1e-6 * abs(elbo[1])     # VB-EM is monotone
    assert int((m.W.info != 0).sum()) == 0
    m.update_assignments(Xc, Yc)
    rs = m.p.sum(-1)
    assert_maxabs(rs, torch.ones_like(rs), 2e-5, "rows of p sum to 1")
    assert_close(m.NA, m.p.double().sum(0), 1e-5, "NA vs p.sum")
    assert abs(float(m.NA.double().sum()) - N) < 1e-6 * N
    assert torch.isfinite(m.logZ).all() and m.logZ.shape == (N,)
    # chunk independence of the E-step (bitwise)
    p_full, lz_full = m.p.clone(), m.logZ.clone()
    cut = N // 2 + 4321
    m.update_assignments(Xc[:cut], Yc[:cut])
    assert torch.equal(m.p, p_full[:cut]) and torch.equal(m.logZ, lz_full[:cut])
    # M-step: the posterior mean of the busiest component from fp64 sums over the SAME responsibilities
    # (transforms/MatrixNormalWishart.py:105-108: invV = invV_0 + SExx, mu = (mu_0 invV_0 + SEyx) invV^-1)
    k0 = int(p_full.sum(0).argmax())
    w = p_full[:, k0].double()
    SExx = torch.zeros(p + 1, p + 1, dtype=torch.float64, device=DEV)
    SEyx = torch.zeros(n, p + 1, dtype=torch.float64, device=DEV)
    for a in range(0, N, 1 << 19):
        xa = torch.cat([X[a:a + (1 << 19)], torch.ones(min(1 << 19, N - a), 1, device=DEV)], -1).double()
        wa = w[a:a + (1 << 19)].unsqueeze(-1)
        SExx += (xa * wa).t() @ xa
        SEyx += (Y[a:a + (1 << 19)].double() * wa).t() @ xa
    W0 = m.W
    invV0, mu0 = W0.invV_0[k0].double(), W0.mu_0[k0].double()
    m.p = p_full
    m.NA = p_full.sum(0)
    W0.raw_update(X.view(N, 1, p, 1), Y.view(N, 1, n, 1), p=p_full, lr=1.0)
    invV = invV0 + SExx
    mu = torch.linalg.solve(invV, (mu0 @ invV0 + SEyx).t()).t()
    assert_close(W0.invV[k0], invV, 1e-5, "invV of the busiest component vs fp64 sums")
    assert_close(W0.mu[k0], mu, 1e-4, "mu of the busiest component vs fp64 sums")


def test_cfg4_full_size_properties():
    """BASELINE.json configs[3] at its full size (ARHMM, 4096 sequences x T = 1024, d = 16, K = 32): the forward-backward kernel
    against the fp64 restatement on a handful of full-length sequences, and reductions that must hold exactly."""
    import numpy as np
    S, T, d, K = 4096, 1024, 16, 32
    g = torch.Generator(device=DEV).manual_seed(2)
    A = 0.95 * torch.linalg.qr(torch.randn(K, d, d, generator=g, device=DEV))[0]
    P = 4 * torch.eye(K, device=DEV) + torch.rand(K, K, generator=g, device=DEV)
    P = P / P.sum(-1, keepdim=True)
    y = torch.zeros(T + 1, S, d, device=DEV)
    zt = torch.randint(K, (S,), generator=g, device=DEV)
    y[0] = torch.randn(S, d, generator=g, device=DEV)
    for t in range(T):
        y[t + 1] = torch.einsum("sij,sj->si", A[zt], y[t]) + 0.3 * torch.randn(S, d, generator=g, device=DEV)
        zt = torch.multinomial(P[zt], 1, generator=g).squeeze(-1)
    X = y[:-1].reshape(T, S, 1, d, 1).contiguous()
    Y = y[1:].reshape(T, S, 1, d, 1).contiguous()
    torch.manual_seed(0)
    h = V.ARHMM(K, d, d).to(DEV)
    elbo = []
    for _ in range(3):
        h.update((X, Y), iters=1, lr=1)
        elbo.append(float(h.ELBO_last))
    assert all(np.isfinite(elbo)) and elbo[2] > elbo[0]
    # one more E-step by hand: emission logits -> forward-backward
    ol = h.obs_logits((X, Y))
    assert ol.shape == (T, S, K)
    p, SEzz, SEz0, logZ = h.forward_backward_logits(ol.clone())
    assert p.shape == (T, S, K) and SEzz.shape == (S, K, K) and SEz0.shape == (S, K) and logZ.shape == (S,)
    rs = p.sum(-1)
    assert_maxabs(rs, torch.ones_like(rs), 2e-5, "smoothed marginals sum to 1")
    # every step (T - 1 transitions + the initial one) adds a normalised joint to SEzz; SEz0 is a distribution
    assert_maxabs(SEzz.double().sum((-1, -2)), torch.full((S,), float(T), dtype=torch.float64), 2e-5 * T, "sum of SEzz per sequence")
    assert_maxabs(SEz0.sum(-1), torch.ones(S), 2e-5, "SEz0 sums to 1")
    # the column sums of SEzz are the smoothed marginals summed over time (xi_t marginalised over the previous state)
    assert_close(SEzz.double().sum(-2), p.double().sum(0), 2e-5, "SEzz column sums vs sum_t p_t")
    # fp64 restatement of models/HMM.py:72-105 on a few full-length sequences
    idx = torch.tensor([0, 1, 1777, 4095], device=DEV)
    ref = O.hmm_new(None, K, dtype=torch.float64)
    ref["transition"]["alpha"] = h.transition.alpha.detach().cpu().double()
    ref["transition"]["alpha_0"] = h.transition.alpha_0.detach().cpu().double()
    ref["initial"]["alpha"] = h.initial.alpha.detach().cpu().double()
    ref["initial"]["alpha_0"] = h.initial.alpha_0.detach().cpu().double()
    pr, SEzzr, SEz0r, logZr = O.hmm_forward_backward_logits(ref, ol[:, idx].detach().cpu().double())
    L = float(ol[:, idx].abs().max())
    assert_maxabs(p[:, idx].cpu().double(), pr, max(5e-5, 1e-6 * L), f"p vs fp64 forward-backward (|logit| {L:.1e})")
    assert_close(logZ[idx], logZr, 1e-5, "logZ per sequence")
    assert_close(SEzz[idx], SEzzr, 1e-4, "SEzz per sequence")
    assert_maxabs(SEz0[idx].cpu().double(), SEz0r, 5e-5, "SEz0")


# ---- degenerate layouts of the matrix-normal models and the HMM glue (step-wise against the fp64 oracle) -------------------

@pytest.mark.parametrize("N,n,p,K,pad", [(600, 1, 1, 2, True), (3000, 1, 3, 1, True), (2500, 5, 1, 3, False), (4000, 2, 2, 1, False),
                                         (9, 3, 2, 2, True)])
def test_molt_degenerate_shapes(N, n, p, K, pad):
    """MixtureofLinearTransforms with one output, one regressor, one expert, with and without the padded column."""
    g = torch.Generator().manual_seed(11 * N + n + p + K)
    X = torch.randn(N, p, 1, generator=g)
    Wt = torch.randn(K, n, p, generator=g)
    z = torch.randint(K, (N,), generator=g)
    Y = (torch.einsum("nij,nj->ni", Wt[z], X[..., 0]) + 0.2 * torch.randn(N, n, generator=g)).unsqueeze(-1)
    torch.manual_seed(2)
    m = V.MixtureofLinearTransforms(n, p, K, pad_X=pad)
    ref = O.molt_new(n, p, K, pad_X=pad)
    O.load_state(ref, {"W.mu": m.W.mu.clone(), "pi.alpha": m.pi.alpha.clone()})
    O.to_dtype(ref, torch.float64)
    m.to(DEV)
    Xd, Yd = X.to(DEV), Y.to(DEV)
    for it in range(3):
        set_state(m, {k: v.float() for k, v in O.flatten_state(ref).items()})
        m.raw_update(Xd, Yd, iters=1)
        tr = O.molt_raw_update(ref, X.double(), Y.double(), 1, exact=True)
        assert m.p.shape == (N, K) and m.logZ.shape == (N,)
        assert abs(float(m.ELBO_last) - float(tr[0])) <= PARITY * abs(float(tr[0])), (it, float(m.ELBO_last), float(tr[0]))
        _p_gate(m.p, ref, f"MoLT N={N} n={n} p={p} K={K} pad={pad} it{it}")
        assert_close(m.logZ, ref["logZ"], PARITY, f"logZ_n it{it}")
        flat = O.flatten_state(ref)
        for k in MOLT_STATE:
            assert_close(get(m, k), flat[k], 3e-4 if it == 0 else PARITY, f"{k} it{it}")


@pytest.mark.parametrize("K,n,T,S", [(1, 2, 12, 5), (2, 3, 1, 4), (3, 2, 7, 1), (4, 1, 30, 6)])
def test_arhmm_degenerate_shapes(K, n, T, S):
    """ARHMM with one state, one time step, one sequence, one output dimension."""
    g = torch.Generator().manual_seed(5 * K + n + T + S)
    X = torch.randn(T, S, 1, n, 1, generator=g)
    Y = (0.7 * X + 0.3 * torch.randn(T, S, 1, n, 1, generator=g))
    torch.manual_seed(9)
    h = V.ARHMM(K, n, n)
    ref = O.arhmm_new(K, n, n)
    O.load_state(ref, {"obs.mu": h.obs_dist.mu.clone(), "transition.alpha": h.transition.alpha.clone(),
                       "initial.alpha": h.initial.alpha.clone()})
    O.to_dtype(ref, torch.float64)
    h.to(DEV)
    Xd, Yd = X.to(DEV), Y.to(DEV)
    keys = ("obs.mu", "obs.invV", "obs.V", "obs.invU.invU", "obs.invU.U", "obs.invU.nu", "transition.alpha", "initial.alpha")
    for it in range(3):
        set_state(h, {k.replace("obs.", "obs_dist."): v.float() for k, v in O.flatten_state(ref).items()})
        h.update((Xd, Yd), iters=1)
        tr = O.arhmm_update(ref, X.double(), Y.double(), 1, exact=True)
        assert h.p.shape == (T, S, K)
        assert_maxabs(h.p.cpu().double(), ref["p"], 5e-5, f"p it{it}")
        assert_close(h.logZ, ref["logZ"], PARITY, f"logZ it{it}")
        assert_close(h.NA, ref["NA"], PARITY, f"NA it{it}")
        assert abs(float(h.ELBO_last) - float(tr[0])) <= PARITY * abs(float(tr[0])), it
        flat = O.flatten_state(ref)
        for k in keys:
            assert_close(get(h, k.replace("obs.", "obs_dist.")), flat[k], 3e-4 if it == 0 else PARITY, f"{k} it{it}")


def test_hmm_more_than_32_states():
    """K = 40 hidden states: beyond the forward-backward kernel's one-warp-per-sequence layout, so the recursion runs as batched
    torch ops on the device around the same K1 / K2 / K3 / K5 kernels (hmm.py) — against the fp64 oracle, step-wise."""
    K, d, T, S = 40, 3, 25, 12
    g = torch.Generator().manual_seed(77)
    cent = 3.0 * torch.randn(K, d, generator=g)
    y = cent[torch.randint(K, (T, S), generator=g)] + 0.4 * torch.randn(T, S, d, generator=g)
    torch.manual_seed(6)
    h = V.HMM(V.NormalInverseWishart(event_shape=(d,), batch_shape=(K,)))
    ref = O.hmm_new(O.niw_new((d,), (K,)), K)
    O.load_state(ref, {"obs.mu": h.obs_dist.mu.clone(), "transition.alpha": h.transition.alpha.clone(),
                       "initial.alpha": h.initial.alpha.clone()})
    O.to_dtype(ref, torch.float64)
    h.to(DEV)
    yd = y.to(DEV)
    keys = ("obs.mu", "obs.lambda_mu", "obs.invU.invU", "obs.invU.U", "obs.invU.nu", "transition.alpha", "initial.alpha")
    for it in range(3):
        set_state(h, {k.replace("obs.", "obs_dist."): v.float() for k, v in O.flatten_state(ref).items()})
        h.update(yd, iters=1)
        tr = O.hmm_niw_update(ref, y.double(), 1)
        assert h.p.shape == (T, S, K)
        assert_maxabs(h.p.cpu().double(), ref["p"], 5e-5, f"p it{it}")
        assert_close(h.logZ, ref["logZ"], PARITY, f"logZ it{it}")
        assert abs(float(h.ELBO_last) - float(tr[0])) <= PARITY * abs(float(tr[0])), it
        flat = O.flatten_state(ref)
        for k in keys:
            assert_close(get(h, k.replace("obs.", "obs_dist.")), flat[k], 3e-4 if it == 0 else PARITY, f"{k} it{it}")


def test_views_and_offsets_give_the_same_bits():
    """Rows handed over as a non-contiguous view, as a slice at an odd element offset (not 16-byte aligned) and as fp64 give the
    results of the contiguous fp32 copy bit for bit (the binding re-bases / converts; the kernels see the same values)."""
    N, d, K = 3000, 16, 8
    g = torch.Generator().manual_seed(3)
    base = torch.randn(N + 1, 2 * d + 1, generator=g).to(DEV)
    variants = {"contiguous": base[1:, 1:2 * d + 1:2].contiguous(), "strided view": base[1:, 1:2 * d + 1:2],
                "odd offset": base.reshape(-1)[1:1 + N * d].view(N, d)}
    variants["fp64"] = variants["contiguous"].double()
    outs = {}
    for name, X in variants.items():
        Xr = variants["contiguous"] if name in ("strided view", "fp64") else X
        torch.manual_seed(1)
        m = V.GaussianMixtureModel(K, d)
        m.dist.mu = variants["contiguous"][:K].clone().cpu() if name != "odd offset" else X[:K].float().clone().cpu()
        m.to(DEV)
        m.update(X, 2)
        outs[name] = (m.p.clone(), m.dist.mu.clone(), m.ELBO_last.clone(), Xr)
    for name in ("strided view", "fp64"):
        for a, b in zip(outs[name][:3], outs["contiguous"][:3]):
            assert torch.equal(a, b), name
    # the odd-offset slice holds different numbers; it must agree with ITS contiguous clone
    Xo = variants["odd offset"].clone()
    torch.manual_seed(1)
    m = V.GaussianMixtureModel(K, d)
    m.dist.mu = Xo[:K].clone().cpu()
    m.to(DEV)
    m.update(Xo, 2)
    for a, b in zip(outs["odd offset"][:3], (m.p, m.dist.mu, m.ELBO_last)):
        assert torch.equal(a, b), "odd offset"


def test_interleaved_models_and_changing_inputs():
    """The binding caches three things between calls — the grow-only workspace, the E-step's weight images for the next Gram and
    the transposed sample image of X — all keyed on tensor identity / version.  Usage patterns that would expose a stale cache:
    two models updated alternately on the same stream (one of them on a second data set of another size), rows edited in
    place between iterations, and a model that returns to an earlier, smaller data set.  Every result must equal, bit for bit,
    the result of the same model run on its own in a fresh state."""
    from pyvbmp_b200 import _lib
    d, K = 64, 32
    g = torch.Generator(device=DEV).manual_seed(12)
    XA = torch.randn(6000, d, generator=g, device=DEV) * 1.3
    XB = torch.randn(9001, d, generator=g, device=DEV) * 0.7 + 0.5

    def fresh(seed, X):
        torch.manual_seed(seed)
        m = V.GaussianMixtureModel(K, d)
        m.dist.mu = X[:K].clone().cpu()
        return m.to(DEV)

    def state(m):
        return [t.clone() for t in (m.dist.mu, m.dist.invU.invU, m.dist.lambda_mu, m.pi.alpha, m.ELBO_last, m.p)]

    def same(a, b, what):
        for i, (x, y) in enumerate(zip(a, b)):
            assert torch.equal(x, y), (what, i)

    # on their own
    a = fresh(1, XA)
    for _ in range(3):
        a.update(XA, 1)
    alone_a = state(a)
    b = fresh(2, XB)
    for _ in range(3):
        b.update(XB, 1)
    alone_b = state(b)
    # alternately (and with the E-step and the M-step of the two models interleaved by hand in the last round)
    _lib.release_workspaces()
    a, b = fresh(1, XA), fresh(2, XB)
    for _ in range(2):
        a.update(XA, 1)
        b.update(XB, 1)
    a.update_assignments(XA)
    b.update_assignments(XB)                       # B's E-step lands between A's E-step and A's M-step
    ea = a.ELBO()
    a.update_parms(XA, 1.0)
    a.ELBO_last = ea
    eb = b.ELBO()
    b.update_parms(XB, 1.0)
    b.ELBO_last = eb
    same(state(a), alone_a, "model A interleaved")
    same(state(b), alone_b, "model B interleaved")
    # rows edited in place between iterations: the second iteration must see the new rows
    Xe = XA.clone()
    c = fresh(3, Xe)
    c.update(Xe, 1)
    Xe.mul_(1.5)
    c.update(Xe, 1)
    ref = fresh(3, XA)
    ref.update(XA.clone(), 1)
    ref.update((XA * 1.5).contiguous(), 1)
    same(state(c), state(ref), "rows edited in place")
    # back to an earlier, smaller data set after a larger one
    e = fresh(4, XA)
    e.update(XA, 1)
    e.update(XB, 1)
    e.update(XA, 1)
    f = fresh(4, XA)
    f.update(XA.clone(), 1)
    _lib.release_workspaces()
    f.update(XB.clone(), 1)
    _lib.release_workspaces()
    f.update(XA.clone(), 1)
    same(state(e), state(f), "returning to a smaller data set")


@pytest.mark.parametrize("N,n,p,K,pad", [(5000, 32, 32, 64, True), (3000, 16, 8, 10, True), (100, 5, 3, 3, False), (70000, 32, 16, 128, True),
                                         (2049, 16, 16, 4, False), (4000, 1, 1, 2, True), (6000, 12, 7, 36, True)])
def test_molt_predict_vs_fp64_oracle(N, n, p, K, pad):
    """MixtureofLinearTransforms.predict at model level across the windows of its kernels (tcgen05 row GEMM with K <= 64,
    the warp-level one otherwise; the SYRK moments kernel at n = 16 / 32, the register-tiled and row-per-lane ones otherwise;
    K % 4 != 0; a few rows): predictive mean, covariance and gate probabilities against the fp64 oracle on a fitted model."""
    g = torch.Generator().manual_seed(13 * N + n + p + K)
    X = torch.randn(N, p, 1, generator=g)
    Wt = torch.randn(K, n, p, generator=g) / p ** 0.5
    z = torch.randint(K, (N,), generator=g)
    Y = (torch.einsum("nij,nj->ni", Wt[z], X[..., 0]) + 0.1 * torch.randn(N, n, generator=g)).unsqueeze(-1)
    torch.manual_seed(8)
    m = V.MixtureofLinearTransforms(n, p, K, pad_X=pad)
    ref = O.molt_new(n, p, K, pad_X=pad)
    O.load_state(ref, {"W.mu": m.W.mu.clone(), "pi.alpha": m.pi.alpha.clone()})
    O.to_dtype(ref, torch.float64)
    O.molt_raw_update(ref, X.double(), Y.double(), 3, exact=False, chunk=8192)         # a fitted state, from the oracle
    m.to(DEV)
    set_state(m, {k: v.float() for k, v in O.flatten_state(ref).items()})
    Xq = torch.randn(N, p, 1, generator=g)
    mu_r, Sig_r, p_r = O.molt_predict(ref, Xq.double())
    pY, pr = m.predict(Xq.to(DEV))
    mu, Sig = pY.mean(), pY.ESigma()
    assert mu.shape == (N, n, 1) and Sig.shape == (N, n, n) and pr.shape == (N, K)
    assert_maxabs(pr.cpu().double(), p_r, 5e-5, "gate probabilities")
    assert_close(mu, mu_r, 2e-5, "predictive mean")
    # covariance entries relative to the largest one of the same sample (mu mu^T is subtracted from a sum of the same size)
    scale = Sig_r.abs().amax((-1, -2), keepdim=True)
    err = float(((Sig.cpu().double() - Sig_r).abs() / scale).max())
    assert err < 5e-5, f"predictive covariance: {err:.2e}"
    assert torch.equal(Sig, Sig.transpose(-1, -2)) or float((Sig - Sig.transpose(-1, -2)).abs().max()) <= 1e-6 * float(scale.max())


@pytest.mark.parametrize("N,n,p,K,shared", [(4096, 32, 32, 64, False), (100, 5, 3, 3, False), (3000, 8, 4, 6, False), (5000, 16, 16, 8, True)])
def test_molt_update_given_beliefs_windows(N, n, p, K, shared):
    """update(pX, pY) inside (N >= 2048, K % 4 == 0) and outside the windows of vbmp_rowterm / vbmp_wsum, and with ONE covariance
    shared by all samples (kept as a single row: its weighted sum is NA_k Sigma).  Two step-wise iterations, fp64 oracle."""
    X, Y = _cfg3_data(N, n, p, K, seed=31)
    g = torch.Generator().manual_seed(32)
    if shared:
        Ax, Ay = 0.2 * torch.randn(p, p, generator=g), 0.2 * torch.randn(n, n, generator=g)
        Sx1, Sy1 = Ax @ Ax.t() + 0.05 * torch.eye(p), Ay @ Ay.t() + 0.05 * torch.eye(n)
        Sx, Sy = Sx1.expand(N, p, p), Sy1.expand(N, n, n)
        Sx_in, Sy_in = Sx1, Sy1
    else:
        Ax, Ay = 0.2 * torch.randn(N, p, p, generator=g), 0.2 * torch.randn(N, n, n, generator=g)
        Sx = Ax @ Ax.transpose(-1, -2) + 0.05 * torch.eye(p)
        Sy = Ay @ Ay.transpose(-1, -2) + 0.05 * torch.eye(n)
        Sx_in, Sy_in = Sx, Sy
    torch.manual_seed(6)
    m = V.MixtureofLinearTransforms(n, p, K)
    ref = O.molt_new(n, p, K)
    O.load_state(ref, {"W.mu": m.W.mu.clone(), "pi.alpha": m.pi.alpha.clone()})
    O.to_dtype(ref, torch.float64)
    m.to(DEV)
    pX = V.MultivariateNormal_vector_format(mu=X.to(DEV), Sigma=Sx_in.to(DEV))
    pY = V.MultivariateNormal_vector_format(mu=Y.to(DEV), Sigma=Sy_in.to(DEV))
    for it in range(2):
        state_in = {k: v.clone() for k, v in O.flatten_state(ref).items()}
        set_state(m, {k: v.float() for k, v in state_in.items()})
        m.update(pX, pY, iters=1)
        elbo = O.molt_update_given(ref, X.double(), Sx.double(), Y.double(), Sy.double())
        assert abs(float(m.ELBO_last) - float(elbo)) <= PARITY * abs(float(elbo)), it
        # the same step in fp32 on the CPU, from the same state: what plain fp32 arithmetic is from fp64 here
        r32 = O.molt_new(n, p, K)
        O.load_state(r32, {k: v.float() for k, v in state_in.items()})
        O.molt_update_given(r32, X, Sx.float(), Y, Sy.float())
        e32 = float((r32["p"].double() - ref["p"]).abs().max())
        print(f"[fp32 CPU restatement] it{it}: max |dp| = {e32:.2e}")
        _p_gate(m.p, ref, f"update(pX, pY) N={N} n={n} p={p} K={K} shared={shared} it{it}")
        assert_close(m.logZ, ref["logZ"], PARITY, f"logZ_n it{it}")
        flat = O.flatten_state(ref)
        for k in MOLT_STATE:
            assert_close(get(m, k), flat[k], 3e-4 if it == 0 else PARITY, f"{k} it{it}")
