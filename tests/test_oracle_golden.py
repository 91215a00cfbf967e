"""Pin the CPU oracle (oracle/vbem_oracle.py) to outputs of the UNMODIFIED reference.

The fixtures were produced by tests/golden/make_golden.py (which imports /root/reference).  The
``exact`` oracle keeps the reference's fp32 op order, so agreement is at fp32 rounding level; the
``fast`` (matmul, fp64) restatement is held to the 1e-4 parity tolerance of BASELINE.json.
"""
import numpy as np
import pytest
import torch

from oracle import vbem_oracle as O
from _util import load_golden, tag, relerr, assert_close

TIGHT = 2e-5     # fp32 re-association noise between two CPU evaluations of the same formulae
PARITY = 1e-4    # BASELINE.json north_star tolerance (ELBO and posterior parameters)


def _gmm_from(fix, t, nc, d, dtype=torch.float32, batch=None, event=None):
    torch.manual_seed(0)
    if batch is None:
        m = O.gmm_new(nc, d)
    else:
        m = O.mixture_new(O.niw_new(event, batch), (nc,))
    O.load_state(m, tag(fix, t))
    if dtype != torch.float32:
        O.to_dtype(m, dtype)
    return m


@pytest.mark.parametrize("name", ["gmm_d2_k6", "gmm_d16_k8_lr05", "gmm_d64_k32_overlap", "gmm_moons_k20"])
def test_gmm_exact_trajectory(name):
    fix = load_golden(name)
    X = torch.as_tensor(fix["X"])
    nc, iters, lr = int(fix["nc"]), int(fix["iters"]), float(fix["lr"])
    m = _gmm_from(fix, "init", nc, X.shape[-1])
    chunk = 256 if X.shape[-1] >= 64 else None
    trace = O.mixture_update(m, X, iters=1, lr=lr, chunk=chunk)
    it1 = tag(fix, "iter1")
    assert_close(m["logZ"], it1["logZ"], TIGHT, "logZ iter1")
    assert_close(m["NA"], it1["NA"], TIGHT, "NA iter1")
    for k in ("dist.mu", "dist.lambda_mu", "dist.invU.invU", "dist.invU.U", "dist.invU.nu",
              "dist.invU.logdet_invU", "pi.alpha"):
        assert_close(O.flatten_state(m)[k], it1[k], 5e-5, k + " iter1")
    assert_close(O.mixture_kl(m), it1["KL"], TIGHT, "KL iter1")
    if "p" in it1:
        assert float((m["p"] - it1["p"]).abs().max()) < 1e-5
    trace += O.mixture_update(m, X, iters=iters - 1, lr=lr, chunk=chunk)
    ref_elbo = fix["ELBO"]
    got = np.array([float(e) for e in trace])
    assert np.max(np.abs(got - ref_elbo) / np.abs(ref_elbo)) < PARITY
    fin = tag(fix, "final")
    assert (m["p"].argmax(-1).numpy() == fix["final/assignment"]).mean() > 0.999
    assert_close(m["dist"]["mu"], fin["dist.mu"], 5e-3, "mu final (free-running, SURVEY F.3)")


@pytest.mark.parametrize("name", ["gmm_d16_k8_lr05", "gmm_d64_k32_overlap"])
def test_gmm_fast_fp64_matches_reference(name):
    """The matmul restatement (ground truth in fp64) reproduces the reference's step at 1e-4."""
    fix = load_golden(name)
    X = torch.as_tensor(fix["X"])
    nc, lr = int(fix["nc"]), float(fix["lr"])
    m = _gmm_from(fix, "init", nc, X.shape[-1], dtype=torch.float64)
    trace = O.mixture_update(m, X.double(), iters=1, lr=lr, exact=False)
    it1 = tag(fix, "iter1")
    assert abs(float(trace[0]) - fix["ELBO"][0]) / abs(fix["ELBO"][0]) < PARITY
    for k in ("dist.mu", "dist.lambda_mu", "dist.invU.invU", "dist.invU.U", "dist.invU.nu", "pi.alpha"):
        assert_close(O.flatten_state(m)[k], it1[k], PARITY, k)
    assert float((m["dist"]["invU"]["logdet_invU"] - it1["dist.invU.logdet_invU"]).abs().max()) < 1e-3
    # the reference's fp32 logits carry ~1e-4 abs noise at iteration 1 (|logit| ~ 1e3 under the broad prior),
    # so its responsibilities sit up to a few 1e-4 from the fp64 truth
    assert float((m["p"] - it1["p"].double()).abs().max()) < 1e-3
    assert (m["p"].argmax(-1) == it1["p"].argmax(-1)).all()


def test_niw_beta_lr_steps():
    fix = load_golden("niw_beta_lr")
    torch.manual_seed(0)
    s = O.niw_new((3,), (4,), scale=0.7)
    O.load_state(s, tag(fix, "init"))
    for i in range(3):
        O.niw_raw_update_exact(s, torch.as_tensor(fix[f"X{i}"]), torch.as_tensor(fix[f"p{i}"]), lr=0.6, beta=0.9)
        ref = tag(fix, f"step{i}")
        flat = O.flatten_state(s)
        for k in ("mu", "lambda_mu", "invU.invU", "invU.U", "invU.nu", "invU.logdet_invU", "SExx", "SEx", "N"):
            assert_close(flat[k], ref[k], TIGHT, f"{k} step{i}")
    assert_close(O.niw_kl(s), fix["final/KL"], TIGHT, "KL")
    assert_close(O.niw_elog_like_exact(s, torch.as_tensor(fix["X2"])), fix["final/Elog_like"], TIGHT, "Elog_like")


def test_niw_fixed_precision_pnone():
    fix = load_golden("niw_fixed_precision_pnone")
    torch.manual_seed(0)
    s = O.niw_new((3,), (2,), fixed_precision=True)
    O.load_state(s, tag(fix, "init"))
    O.niw_raw_update_exact(s, torch.as_tensor(fix["X"]), None, lr=1.0, beta=None)
    ref = tag(fix, "final")
    flat = O.flatten_state(s)
    for k in ("mu", "lambda_mu", "invU.invU", "invU.nu"):
        assert_close(flat[k], ref[k], TIGHT, k)
    assert_close(O.niw_kl(s), fix["final/KL"], TIGHT, "KL")


@pytest.mark.parametrize("name,batch,event,nc,iters", [
    ("mixture_batch3_k6", (3, 6), (2,), 6, 4),
    ("mixture_event32_k5", (5,), (3, 2), 5, 3),
])
def test_mixture_general_shapes(name, batch, event, nc, iters):
    fix = load_golden(name)
    X = torch.as_tensor(fix["X"])
    m = _gmm_from(fix, "init", nc, event[-1], batch=batch, event=event)
    trace = O.mixture_update(m, X, iters=iters)
    got = torch.stack(trace).numpy()
    assert np.max(np.abs(got - fix["ELBO"]) / np.abs(fix["ELBO"])) < PARITY
    fin = tag(fix, "final")
    assert float((m["p"] - fin["p"]).abs().max()) < 1e-4
    assert_close(m["NA"], fin["NA"], 1e-4, "NA")
    assert_close(m["dist"]["mu"], fin["dist.mu"], 1e-3, "mu")
    assert m["dist"]["invU"]["nu"].shape == fin["dist.invU.nu"].shape


def test_niw_hmm_emission_sample_shape_ts():
    fix = load_golden("niw_hmm_emission_ts")
    torch.manual_seed(0)
    s = O.niw_new((2,), (4,))
    O.load_state(s, tag(fix, "init"))
    X, p = torch.as_tensor(fix["X"]).unsqueeze(-2), torch.as_tensor(fix["p"])
    assert_close(O.niw_elog_like_exact(s, X), fix["init/Elog_like"], TIGHT, "Elog_like init")
    O.niw_raw_update_exact(s, X, p)
    assert_close(O.niw_elog_like_exact(s, X), fix["final/Elog_like"], TIGHT, "Elog_like final")
    assert_close(O.niw_kl(s), fix["final/KL"], TIGHT, "KL")


@pytest.mark.parametrize("pad", [1, 0])
def test_mnw_steps(pad):
    fix = load_golden(f"mnw_n4_p5_k3_pad{pad}")
    n, p, K = int(fix["n"]), int(fix["p"]), int(fix["K"])
    torch.manual_seed(0)
    s = O.mnw_new((n, p), (K,), scale=0.8, pad_X=bool(pad))
    O.load_state(s, tag(fix, "init"))
    X, Y, r = (torch.as_tensor(fix[k]) for k in ("X", "Y", "r"))
    assert_close(O.mnw_elog_like_exact(s, X, Y), fix["init/Elog_like"], TIGHT, "Elog_like init")
    assert_close(O.mnw_elog_like_fast(s, X[:, 0, :, 0], Y[:, 0, :, 0]), fix["init/Elog_like"], 5e-5, "fast form")
    assert_close(O.mnw_kl(s), fix["init/KL"], TIGHT, "KL init")
    O.mnw_ss_update(s, *O.mnw_raw_stats_exact(s, X, Y, r), lr=1.0, beta=None)
    ref = tag(fix, "step0")
    flat = O.flatten_state(s)
    for k in ("mu", "invV", "V", "logdetinvV", "invU.invU", "invU.U", "invU.nu", "invU.logdet_invU"):
        assert_close(flat[k], ref[k], 5e-5, k + " step0")
    # the same update fed from the single weighted Gram matrix (the form the CUDA path uses)
    s2 = O.mnw_new((n, p), (K,), scale=0.8, pad_X=bool(pad))
    O.load_state(s2, tag(fix, "init"))
    G = O.weighted_gram_fast(torch.cat([Y[:, 0, :, 0], X[:, 0, :, 0]], -1), r).float()
    O.mnw_ss_update(s2, *O.mnw_gram_blocks(s2, G), lr=1.0, beta=None)
    for k in ("mu", "invV", "V", "invU.invU", "invU.U", "invU.nu"):
        assert_close(O.flatten_state(s2)[k], ref[k], PARITY, k + " via Gram")
    assert_close(O.mnw_elog_like_exact(s, X, Y), fix["step0/Elog_like"], 5e-5, "Elog_like step0")
    assert_close(O.mnw_kl(s), fix["step0/KL"], 5e-5, "KL step0")
    O.mnw_ss_update(s, *O.mnw_raw_stats_exact(s, X, Y, r), lr=0.5, beta=0.8)
    if not pad:
        O.mnw_ss_update(s, *O.mnw_raw_stats_exact(s, X, Y, None), lr=0.5, beta=0.8)
    ref = tag(fix, "step2")
    flat = O.flatten_state(s)
    for k in ("mu", "invV", "V", "invU.invU", "invU.U", "invU.nu"):
        assert_close(flat[k], ref[k], 5e-5, k + " step2")
    assert_close(O.mnw_kl(s), fix["step2/KL"], 5e-5, "KL step2")


@pytest.mark.parametrize("name", ["molt_n3_p4_k5", "molt_n32_p32_k8"])
@pytest.mark.parametrize("exact", [True, False])
def test_molt_trajectory(name, exact):
    fix = load_golden(name)
    n, p, K, iters = (int(fix[k]) for k in ("n", "p", "K", "iters"))
    torch.manual_seed(0)
    m = O.molt_new(n, p, K)
    O.load_state(m, tag(fix, "init"))
    X, Y = torch.as_tensor(fix["X"]).unsqueeze(-1), torch.as_tensor(fix["Y"]).unsqueeze(-1)
    if not exact:
        O.to_dtype(m, torch.float64)
        X, Y = X.double(), Y.double()
    trace = O.molt_raw_update(m, X, Y, iters=1, exact=exact)
    it1 = tag(fix, "iter1")
    assert float((m["p"] - it1["p"]).abs().max()) < (2e-5 if exact else 5e-4)   # fp64 truth vs fp32 reference
    assert_close(m["logZ"], it1["logZ"], 2e-5, "logZ_n")
    for k in ("W.mu", "W.invV", "W.V", "W.invU.invU", "W.invU.U", "W.invU.nu", "pi.alpha"):
        assert_close(O.flatten_state(m)[k], it1[k], PARITY, k)
    trace += O.molt_raw_update(m, X, Y, iters=iters - 1, exact=exact)
    got = np.array([float(e) for e in trace])
    assert np.max(np.abs(got - fix["ELBO"]) / np.abs(fix["ELBO"])) < PARITY
    assert (m["p"].argmax(-1).numpy() == fix["final/assignment"]).mean() > 0.995


@pytest.mark.parametrize("name", ["molt_predict_n3_p4_k5", "molt_predict_n16_p32_k8"])
@pytest.mark.parametrize("f64", [False, True])
def test_molt_predict(name, f64):
    """MixtureofLinearTransforms.predict (transforms/MixtureofLinearTransforms.py:91-108) on held-out inputs."""
    fix = load_golden(name)
    n, p, K = (int(fix[k]) for k in ("n", "p", "K"))
    m = O.molt_new(n, p, K)
    O.load_state(m, tag(fix, "state"))
    Xt = torch.as_tensor(fix["Xt"]).unsqueeze(-1)
    if f64:
        O.to_dtype(m, torch.float64)
        Xt = Xt.double()
    mu, Sigma, pr = O.molt_predict(m, Xt)
    pf = tag(fix, "predict")
    assert float((pr - pf["p"]).abs().max()) < 5e-5
    assert_close(mu, pf["mu"], PARITY, "predict mean")
    assert_close(Sigma, pf["Sigma"], 5 * PARITY, "predict covariance")       # formed as a difference of second moments


def test_arhmm_prxy_trajectory():
    """ARHMM_prXY (models/ARHMM.py:35-46): HMM forward-backward on expectation-input observation logits."""
    fix = load_golden("arhmm_prxy_k4_n2_p3")
    K, n, p = int(fix["K"]), int(fix["n"]), int(fix["p"])
    h = O.arhmm_new(K, n, p)
    O.load_state(h, tag(fix, "init"))
    t = lambda k: torch.as_tensor(fix[k])                                           # noqa: E731
    mux, Sx, muy, Sy = t("mux"), t("Sx"), t("muy"), t("Sy")
    ol = O.mnw_elog_like_given(h["obs"], mux, Sx + mux @ mux.transpose(-2, -1), muy, Sy + muy @ muy.transpose(-2, -1))
    assert_close(ol, fix["init/obs_logits"], 2e-5, "obs_logits")
    trace = O.arhmm_prxy_update(h, mux, Sx, muy, Sy, iters=1)
    it1 = tag(fix, "iter1")
    assert float((h["p"] - it1["p"]).abs().max()) < 5e-5
    assert_close(h["logZ"], it1["logZ"], PARITY, "logZ")
    assert_close(h["NA"], it1["NA"], PARITY, "NA")
    for k in ("obs.mu", "obs.invV", "obs.invU.invU", "transition.alpha", "initial.alpha"):
        assert_close(O.flatten_state(h)[k], it1[k], PARITY, k)
    trace += O.arhmm_prxy_update(h, mux, Sx, muy, Sy, iters=2)
    got = np.array([float(e) for e in trace])
    assert np.max(np.abs(got - fix["ELBO"]) / np.abs(fix["ELBO"])) < PARITY


def test_arhmm_prxry_trajectory():
    """ARHMM_prXRY (models/ARHMM.py:55-77): latent-regressor belief stacked on observed regressors, observed outputs."""
    fix = load_golden("arhmm_prxry_k4_n2_p21")
    K, n, p1, p2 = (int(fix[k]) for k in ("K", "n", "p1", "p2"))
    h = O.arhmm_prxry_new(K, n, p1, p2)
    O.load_state(h, tag(fix, "init"))
    t = lambda k: torch.as_tensor(fix[k])                                           # noqa: E731
    mux, Sx, R, Y = t("mux"), t("Sx"), t("R"), t("Y")
    ol = O.mnw_elog_like_given(h["obs"], *O.arhmm_prxry_beliefs(mux, Sx, R, Y))
    assert_close(ol, fix["init/obs_logits"], 2e-5, "obs_logits")
    trace = O.arhmm_prxry_update(h, mux, Sx, R, Y, iters=1)
    it1 = tag(fix, "iter1")
    assert float((h["p"] - it1["p"]).abs().max()) < 5e-5
    assert_close(h["logZ"], it1["logZ"], PARITY, "logZ")
    assert_close(h["NA"], it1["NA"], PARITY, "NA")
    for k in ("obs.mu", "obs.invV", "obs.invU.invU", "transition.alpha", "initial.alpha"):
        assert_close(O.flatten_state(h)[k], it1[k], PARITY, k)
    trace += O.arhmm_prxry_update(h, mux, Sx, R, Y, iters=2)
    got = np.array([float(e) for e in trace])
    assert np.max(np.abs(got - fix["ELBO"]) / np.abs(fix["ELBO"])) < PARITY


@pytest.mark.parametrize("name", ["hmm_niw_k6", "hmm_event32_k5", "hmm_masked_k6", "hmm_ptemp2_k6", "hmm_masked_ptemp05_k6"])
def test_hmm_niw_trajectory(name):
    """models.HMM with NormalInverseWishart emissions (models/HMM.py:113-152; tests/test_models.py:293-314, :398-409 layouts)."""
    fix = load_golden(name)
    ev, bs = tuple(int(v) for v in fix["event_shape"]), tuple(int(v) for v in fix["batch_shape"])
    torch.manual_seed(0)
    h = O.hmm_new(O.niw_new(ev, bs), bs[-1])
    O.load_state(h, tag(fix, "init"))              # a transition mask lives in the state: zeros in transition.alpha_0 / alpha
    h["ptemp"] = float(fix["ptemp"])
    y = torch.as_tensor(fix["y"])
    ol = O.niw_elog_like_exact(h["obs"], y.unsqueeze(-1 - len(ev)))
    assert_close(ol, fix["init/obs_logits"], 2e-5, "obs_logits")
    trace = O.hmm_niw_update(h, y, iters=1)
    it1 = tag(fix, "iter1")
    assert float((h["p"] - it1["p"]).abs().max()) < 5e-5
    assert_close(h["logZ"], it1["logZ"], PARITY, "logZ")
    assert_close(h["NA"], it1["NA"], PARITY, "NA")
    for k in ("obs.mu", "obs.lambda_mu", "obs.invU.invU", "obs.invU.nu", "transition.alpha", "initial.alpha"):
        assert_close(O.flatten_state(h)[k], it1[k], PARITY, k)
    trace += O.hmm_niw_update(h, y, iters=2)
    got = np.array([float(e) for e in trace])
    assert np.max(np.abs(got - fix["ELBO"]) / np.abs(fix["ELBO"])) < PARITY


@pytest.mark.parametrize("name", ["molt_given_n3_p4_k5", "molt_given_n8_p16_k6"])
def test_molt_given_beliefs(name):
    """Expectation-input E and M steps (transforms/MatrixNormalWishart.py:143-172, 234-249;
    transforms/MixtureofLinearTransforms.py:62-90)."""
    fix = load_golden(name)
    n, p, K, lr = int(fix["n"]), int(fix["p"]), int(fix["K"]), float(fix["lr"])
    m = O.molt_new(n, p, K)
    O.load_state(m, tag(fix, "state"))
    mux, muy = torch.as_tensor(fix["mux"]).unsqueeze(-1), torch.as_tensor(fix["muy"]).unsqueeze(-1)
    Sx, Sy = torch.as_tensor(fix["Sx"]), torch.as_tensor(fix["Sy"])
    Exx = (Sx + mux @ mux.transpose(-2, -1)).unsqueeze(-3)
    Eyy = (Sy + muy @ muy.transpose(-2, -1)).unsqueeze(-3)
    ELL = O.mnw_elog_like_given(m["W"], mux.unsqueeze(-3), Exx, muy.unsqueeze(-3), Eyy)
    assert_close(ELL, tag(fix, "given")["ELL"], 2e-5, "Elog_like_given_pX_pY")
    elbo = O.molt_update_given(m, mux, Sx, muy, Sy, lr=lr)
    aft = tag(fix, "after")
    assert abs(float(elbo) - float(fix["after/ELBO"])) <= PARITY * abs(float(fix["after/ELBO"]))
    assert float((m["p"] - aft["p"]).abs().max()) < 5e-5
    assert_close(m["logZ"], aft["logZ"], 2e-5, "logZ_n")
    for k in ("W.mu", "W.invV", "W.V", "W.invU.invU", "W.invU.U", "W.invU.nu", "pi.alpha"):
        assert_close(O.flatten_state(m)[k], aft[k], PARITY, k)


@pytest.mark.parametrize("exact", [True, False])
def test_arhmm_trajectory(exact):
    fix = load_golden("arhmm_k4_n2_p3")
    K, n, p = int(fix["K"]), int(fix["n"]), int(fix["p"])
    torch.manual_seed(0)
    h = O.arhmm_new(K, n, p)
    O.load_state(h, tag(fix, "init"))
    X, Y = torch.as_tensor(fix["X"]), torch.as_tensor(fix["Y"])
    assert_close(O.arhmm_obs_logits(h, X, Y, exact), fix["init/obs_logits"], 5e-5, "obs_logits")
    trace = O.arhmm_update(h, X, Y, iters=1, exact=exact)
    it1 = tag(fix, "iter1")
    assert float((h["p"] - it1["p"]).abs().max()) < (5e-5 if exact else 5e-4)
    assert_close(h["logZ"], it1["logZ"], 2e-5, "logZ")
    assert_close(h["NA"], it1["NA"], 5e-5, "NA")
    for k in ("obs.mu", "obs.invV", "obs.V", "obs.invU.invU", "obs.invU.U", "transition.alpha", "initial.alpha"):
        assert_close(O.flatten_state(h)[k], it1[k], PARITY, k)
    trace += O.arhmm_update(h, X, Y, iters=3, exact=exact)
    got = np.array([float(e) for e in trace])
    assert np.max(np.abs(got - fix["ELBO"]) / np.abs(fix["ELBO"])) < PARITY
