"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (it imports /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Each fixture is a flat .npz: seeded inputs, the reference's initial state (after construction /
``initialize``), and the reference's outputs (per-iteration ELBO, responsibilities or their
argmax, and the posterior state).  State keys follow ``oracle.vbem_oracle.flatten_state`` naming
(``dist.mu``, ``dist.invU.invU`` ...), so the same file feeds the oracle tests (CPU) and the CUDA
parity tests (GPU).  The reference has no golden vectors of its own; these
files are what pins the oracle.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("PYVBMP_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
import dists  # noqa: E402
import transforms  # noqa: E402
import models  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def T(x):
    return x.detach().clone().contiguous().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def wishart_state(w, pre):
    return {pre + k: T(getattr(w, k)) for k in ("invU_0", "nu_0", "logdet_invU_0", "invU", "U", "nu", "logdet_invU")}


def niw_state(d, pre):
    out = {pre + k: T(getattr(d, k)) for k in ("lambda_mu_0", "lambda_mu", "mu_0", "mu")}
    out.update(wishart_state(d.invU, pre + "invU."))
    return out


def mnw_state(d, pre):
    out = {pre + k: T(getattr(d, k)) for k in ("mu_0", "mu", "invV_0", "invV", "V", "logdetinvV", "logdetinvV_0")}
    out.update(wishart_state(d.invU, pre + "invU."))
    return out


def dir_state(d, pre):
    return {pre + "alpha_0": T(d.alpha_0), pre + "alpha": T(d.alpha)}


def tagged(state, tag):
    return {tag + "/" + k: v for k, v in state.items()}


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, {len(arrs)} arrays")


def gmm_state(m):
    s = niw_state(m.dist, "dist.")
    s.update(dir_state(m.pi, "pi."))
    return s


def two_moons(n_per, seed):
    """examples/two_moons.py:4-21 data recipe."""
    g = torch.Generator().manual_seed(seed)
    x = torch.linspace(-np.pi / 2, np.pi / 2, n_per)
    a = torch.stack([torch.sin(x), torch.cos(x) - 0.25], -1)
    b = torch.stack([torch.sin(x) + 1.0, -torch.cos(x) + 0.25], -1)
    X = torch.cat([a, b], 0)
    X = X + 0.05 * torch.randn(X.shape, generator=g)
    return X / X.std()


def run_gmm(name, X, nc, iters, lr=1.0, seed=0, keep_p=True):
    torch.manual_seed(seed)
    m = models.GaussianMixtureModel(nc, X.shape[-1])
    m.initialize(X)
    out = {"X": T(X), "nc": nc, "iters": iters, "lr": lr}
    out.update(tagged(gmm_state(m), "init"))
    elbo = []
    for i in range(iters):
        m.update(X, 1, lr)
        elbo.append(float(m.ELBO_last))
        if i == 0:
            out.update(tagged(gmm_state(m), "iter1"))
            out["iter1/logZ"] = T(m.logZ)
            out["iter1/NA"] = T(m.NA)
            out["iter1/KL"] = T(m.KLqprior())
            if keep_p:
                out["iter1/p"] = T(m.p)
    out.update(tagged(gmm_state(m), "final"))
    out["final/logZ"] = T(m.logZ)
    out["final/NA"] = T(m.NA)
    out["final/assignment"] = T(m.assignment()).astype(np.int32)
    if keep_p:
        out["final/p"] = T(m.p)
    # one more E-step on the final parameters: logits straight from dist.Elog_like
    out["final/Elog_like"] = T(m.Elog_like(X)) if keep_p else T(m.Elog_like(X[:256]))
    out["final/KL"] = T(m.KLqprior())
    out["ELBO"] = np.asarray(elbo, dtype=np.float64)
    save(name, **out)


def gen_gmm():
    g = torch.Generator().manual_seed(11)
    # tests/test_dists.py:202-220 recipe: 6 blobs in 2-d, N=400
    mu = 4 * torch.randn(6, 2, generator=g)
    z = torch.randint(6, (400,), generator=g)
    X = mu[z] + 0.6 * torch.randn(400, 2, generator=g)
    run_gmm("gmm_d2_k6", X, 6, 8, seed=1)
    # config 1: two-moons, N=10k, d=2, K=20 (p not stored: 10k x 20)
    run_gmm("gmm_moons_k20", two_moons(5000, 5), 20, 20, seed=0, keep_p=False)
    # mid-size full covariance with lr<1
    A = torch.randn(5, 16, 16, generator=g) / 4 + torch.eye(16)
    mu = 2.0 * torch.randn(5, 16, generator=g)
    z = torch.randint(5, (768,), generator=g)
    X = mu[z] + torch.einsum("nij,nj->ni", A[z], torch.randn(768, 16, generator=g))
    run_gmm("gmm_d16_k8_lr05", X, 8, 4, lr=0.5, seed=2)
    # config-2 shaped slice (overlapping clusters variant): d=64, K=32 keeps the file small
    mu = 0.3 * torch.randn(32, 64, generator=g)
    z = torch.randint(32, (1024,), generator=g)
    X = mu[z] + torch.randn(1024, 64, generator=g)
    run_gmm("gmm_d64_k32_overlap", X, 32, 3, seed=3, keep_p=True)


def gen_niw_variants():
    g = torch.Generator().manual_seed(21)
    # raw_update with beta forgetting and lr<1, p given  (NormalInverseWishart.py:49-86)
    torch.manual_seed(4)
    d = dists.NormalInverseWishart(event_shape=(3,), batch_shape=(4,), scale=0.7)
    out = tagged(niw_state(d, ""), "init")
    for i in range(3):
        X = torch.randn(50, 1, 3, generator=g) + i
        p = torch.rand(50, 4, generator=g)
        d.raw_update(X, p, lr=0.6, beta=0.9)
        out[f"X{i}"], out[f"p{i}"] = T(X), T(p)
        out.update(tagged(niw_state(d, ""), f"step{i}"))
        out[f"step{i}/SExx"], out[f"step{i}/SEx"], out[f"step{i}/N"] = T(d.SExx), T(d.SEx), T(d.N)
    out["final/KL"] = T(d.KLqprior())
    out["final/Elog_like"] = T(d.Elog_like(X))
    save("niw_beta_lr", **out)

    # fixed_precision + p=None
    torch.manual_seed(5)
    d = dists.NormalInverseWishart(event_shape=(3,), batch_shape=(2,), fixed_precision=True)
    out = tagged(niw_state(d, ""), "init")
    X = torch.randn(40, 2, 3, generator=g)
    d.raw_update(X, None, lr=1.0, beta=None)
    out["X"] = T(X)
    out.update(tagged(niw_state(d, ""), "final"))
    out["final/KL"] = T(d.KLqprior())
    save("niw_fixed_precision_pnone", **out)

    # Mixture over NIW with a leading replica batch dim: tests/test_dists.py:256-276
    torch.manual_seed(6)
    dist = dists.NormalInverseWishart(event_shape=(2,), batch_shape=(3, 6), scale=0.5)
    m = dists.Mixture(dist, event_shape=(6,))
    X = torch.randn(200, 3, 2, generator=g) * 2
    out = {"X": T(X)}
    out.update(tagged(gmm_state(m), "init"))
    el = []
    for i in range(4):
        m.update(X, 1, 1.0)
        el.append(T(m.ELBO_last))
    out.update(tagged(gmm_state(m), "final"))
    out["final/p"], out["final/NA"], out["final/logZ"] = T(m.p), T(m.NA), T(m.logZ)
    out["ELBO"] = np.stack(el)
    save("mixture_batch3_k6", **out)

    # Mixture with event_dim>1 NIW: tests/test_dists.py:278-282
    torch.manual_seed(7)
    dist = dists.NormalInverseWishart(event_shape=(3, 2), batch_shape=(5,), scale=0.5)
    m = dists.Mixture(dist, event_shape=(5,))
    X = torch.randn(150, 3, 2, generator=g) * 2
    out = {"X": T(X)}
    out.update(tagged(gmm_state(m), "init"))
    el = []
    for i in range(3):
        m.update(X, 1, 1.0)
        el.append(T(m.ELBO_last))
    out.update(tagged(gmm_state(m), "final"))
    out["final/p"], out["final/NA"], out["final/logZ"] = T(m.p), T(m.NA), T(m.logZ)
    out["ELBO"] = np.stack(el)
    save("mixture_event32_k5", **out)

    # NIW as HMM emission, sample shape (T,S): models/HMM.py:113-117, 138-139
    torch.manual_seed(8)
    d = dists.NormalInverseWishart(event_shape=(2,), batch_shape=(4,))
    Xs = torch.randn(12, 9, 2, generator=g) * 1.5
    p = torch.softmax(torch.randn(12, 9, 4, generator=g), -1)
    out = {"X": T(Xs), "p": T(p)}
    out.update(tagged(niw_state(d, ""), "init"))
    out["init/Elog_like"] = T(d.Elog_like(Xs.unsqueeze(-2)))
    d.raw_update(Xs.unsqueeze(-2), p=p, lr=1.0, beta=None)
    out.update(tagged(niw_state(d, ""), "final"))
    out["final/Elog_like"] = T(d.Elog_like(Xs.unsqueeze(-2)))
    out["final/KL"] = T(d.KLqprior())
    save("niw_hmm_emission_ts", **out)


def gen_mnw():
    g = torch.Generator().manual_seed(31)
    for pad in (True, False):
        torch.manual_seed(9)
        n, p, K, N = 4, 5, 3, 300
        d = transforms.MatrixNormalWishart(event_shape=(n, p), batch_shape=(K,), scale=0.8, pad_X=pad)
        out = {"n": n, "p": p, "K": K, "pad_X": int(pad)}
        out.update(tagged(mnw_state(d, ""), "init"))
        X = torch.randn(N, 1, p, 1, generator=g)
        Wt = torch.randn(K, n, p, generator=g)
        z = torch.randint(K, (N,), generator=g)
        Y = (Wt[z] @ X.squeeze(1)).unsqueeze(1) + 0.3 * torch.randn(N, 1, n, 1, generator=g) + 0.5
        r = torch.softmax(torch.randn(N, K, generator=g), -1)
        out["X"], out["Y"], out["r"] = T(X), T(Y), T(r)
        out["init/Elog_like"] = T(d.Elog_like(X, Y))
        out["init/KL"] = T(d.KLqprior())
        d.raw_update(X, Y, p=r, lr=1.0, beta=None)
        out.update(tagged(mnw_state(d, ""), "step0"))
        out["step0/Elog_like"] = T(d.Elog_like(X, Y))
        out["step0/KL"] = T(d.KLqprior())
        d.raw_update(X, Y, p=r, lr=0.5, beta=0.8)
        if not pad:   # the reference's p=None + pad_X branch only accepts X with the full batch shape (:191-202)
            d.raw_update(X, Y, p=None, lr=0.5, beta=0.8)
        out.update(tagged(mnw_state(d, ""), "step2"))
        out["step2/SExx"], out["step2/SEyx"], out["step2/SEyy"], out["step2/N"] = \
            T(d.SExx), T(d.SEyx), T(d.SEyy), T(d.N)
        out["step2/Elog_like"] = T(d.Elog_like(X, Y))
        out["step2/KL"] = T(d.KLqprior())
        save(f"mnw_n4_p5_k3_pad{int(pad)}", **out)


def molt_state(m):
    s = mnw_state(m.W, "W.")
    s.update(dir_state(m.pi, "pi."))
    return s


def gen_molt():
    g = torch.Generator().manual_seed(41)
    for name, n, p, K, N, iters in (("molt_n3_p4_k5", 3, 4, 5, 600, 5), ("molt_n32_p32_k8", 32, 32, 8, 768, 3)):
        torch.manual_seed(10)
        m = transforms.MixtureofLinearTransforms(n, p, K, pad_X=True)
        X = torch.randn(N, p, generator=g)
        Wt = torch.randn(K, n, p, generator=g) / np.sqrt(p)
        b = torch.randn(K, n, generator=g)
        z = torch.randint(K, (N,), generator=g)
        Y = torch.einsum("nij,nj->ni", Wt[z], X) + b[z] + 0.1 * torch.randn(N, n, generator=g)
        out = {"X": T(X), "Y": T(Y), "n": n, "p": p, "K": K, "iters": iters}
        out.update(tagged(molt_state(m), "init"))
        el = []
        for i in range(iters):
            m.raw_update(X.unsqueeze(-1), Y.unsqueeze(-1), iters=1, lr=1.0)
            el.append(float(m.ELBO_last))
            if i == 0:
                out.update(tagged(molt_state(m), "iter1"))
                out["iter1/p"], out["iter1/logZ"] = T(m.p), T(m.logZ)
        out.update(tagged(molt_state(m), "final"))
        out["final/p"], out["final/logZ"] = T(m.p), T(m.logZ)
        out["final/assignment"] = T(m.assignment()).astype(np.int32)
        out["final/KL"] = T(m.KLqprior())
        out["ELBO"] = np.asarray(el, dtype=np.float64)
        save(name, **out)


def gen_molt_predict():
    """MixtureofLinearTransforms.predict on held-out inputs after a few EM iterations."""
    g = torch.Generator().manual_seed(43)
    for name, n, p, K, N, iters in (("molt_predict_n3_p4_k5", 3, 4, 5, 600, 4), ("molt_predict_n16_p32_k8", 16, 32, 8, 768, 3)):
        torch.manual_seed(12)
        m = transforms.MixtureofLinearTransforms(n, p, K, pad_X=True)
        X = torch.randn(N, p, generator=g)
        Wt = torch.randn(K, n, p, generator=g) / np.sqrt(p)
        b = torch.randn(K, n, generator=g)
        z = torch.randint(K, (N,), generator=g)
        Y = torch.einsum("nij,nj->ni", Wt[z], X) + b[z] + 0.1 * torch.randn(N, n, generator=g)
        m.raw_update(X.unsqueeze(-1), Y.unsqueeze(-1), iters=iters, lr=1.0)
        Xt = 1.5 * torch.randn(200, p, generator=g)
        pY, pr = m.predict(Xt.unsqueeze(-1))
        out = {"Xt": T(Xt), "n": n, "p": p, "K": K}
        out.update(tagged(molt_state(m), "state"))
        out["predict/mu"], out["predict/Sigma"], out["predict/p"] = T(pY.mean()), T(pY.ESigma()), T(pr)
        save(name, **out)


def gen_molt_given():
    """Expectation-input E and M steps: MatrixNormalWishart.Elog_like_given_pX_pY / update(pX, pY, p) and
    MixtureofLinearTransforms.update(pX, pY) with Gaussian beliefs about the inputs and outputs."""
    from dists.MultivariateNormal_vector_format import MultivariateNormal_vector_format as MVN
    g = torch.Generator().manual_seed(47)
    for name, n, p, K, N, lr in (("molt_given_n3_p4_k5", 3, 4, 5, 500, 1.0), ("molt_given_n8_p16_k6", 8, 16, 6, 640, 0.5)):
        torch.manual_seed(14)
        m = transforms.MixtureofLinearTransforms(n, p, K, pad_X=True)
        X = torch.randn(N, p, generator=g)
        Wt = torch.randn(K, n, p, generator=g) / np.sqrt(p)
        b = torch.randn(K, n, generator=g)
        z = torch.randint(K, (N,), generator=g)
        Y = torch.einsum("nij,nj->ni", Wt[z], X) + b[z] + 0.1 * torch.randn(N, n, generator=g)
        m.raw_update(X.unsqueeze(-1), Y.unsqueeze(-1), iters=3, lr=1.0)
        Ax = 0.2 * torch.randn(N, p, p, generator=g)
        Ay = 0.1 * torch.randn(N, n, n, generator=g)
        Sx = Ax @ Ax.transpose(-1, -2) + 0.01 * torch.eye(p)
        Sy = Ay @ Ay.transpose(-1, -2) + 0.01 * torch.eye(n)
        out = {"mux": T(X), "Sx": T(Sx), "muy": T(Y), "Sy": T(Sy), "n": n, "p": p, "K": K, "lr": lr}
        out.update(tagged(molt_state(m), "state"))
        pX, pY = MVN(mu=X.unsqueeze(-1), Sigma=Sx), MVN(mu=Y.unsqueeze(-1), Sigma=Sy)
        out["given/ELL"] = T(m.W.Elog_like_given_pX_pY(pX.unsqueeze(-3), pY.unsqueeze(-3)))
        m.update(pX, pY, iters=1, lr=lr)
        out.update(tagged(molt_state(m), "after"))
        out["after/p"], out["after/logZ"], out["after/ELBO"] = T(m.p), T(m.logZ), np.float64(float(m.ELBO_last))
        save(name, **out)


def gen_arhmm():
    g = torch.Generator().manual_seed(51)
    K, n, p, Tn, S = 4, 2, 3, 40, 25
    # tests/test_models.py:13-38 style switching AR data
    Atrue = torch.randn(K, n, p, generator=g) * 0.7
    btrue = torch.randn(K, n, generator=g)
    trans = 4 * torch.eye(K) + torch.rand(K, K, generator=g)
    trans = trans / trans.sum(-1, keepdim=True)
    z = torch.zeros(Tn, S, dtype=torch.long)
    z[0] = torch.randint(K, (S,), generator=g)
    for t in range(1, Tn):
        z[t] = torch.multinomial(trans[z[t - 1]], 1, generator=g).squeeze(-1)
    X = torch.randn(Tn, S, p, generator=g)
    Y = torch.einsum("tsij,tsj->tsi", Atrue[z], X) + btrue[z] + 0.2 * torch.randn(Tn, S, n, generator=g)
    Xr, Yr = X.unsqueeze(-2).unsqueeze(-1), Y.unsqueeze(-2).unsqueeze(-1)      # (T,S,1,p,1), (T,S,1,n,1)
    torch.manual_seed(12)
    m = models.ARHMM(K, n, p)
    out = {"X": T(Xr), "Y": T(Yr), "K": K, "n": n, "p": p}

    def st():
        s = mnw_state(m.obs_dist, "obs.")
        s.update(dir_state(m.transition, "transition."))
        s.update(dir_state(m.initial, "initial."))
        return s
    out.update(tagged(st(), "init"))
    out["init/obs_logits"] = T(m.obs_logits((Xr, Yr)))
    el = []
    for i in range(4):
        m.update((Xr, Yr), iters=1, lr=1.0)
        el.append(float(m.ELBO_last))
        if i == 0:
            out.update(tagged(st(), "iter1"))
            out["iter1/p"], out["iter1/logZ"], out["iter1/NA"] = T(m.p), T(m.logZ), T(m.NA)
    out.update(tagged(st(), "final"))
    out["final/p"], out["final/logZ"] = T(m.p), T(m.logZ)
    out["ELBO"] = np.asarray(el, dtype=np.float64)
    save("arhmm_k4_n2_p3", **out)


def gen_hmm_variants():
    """models.HMM with NIW emissions in the three layouts of the reference's own script (tests/test_models.py:293-314 plain,
    :353-356 a batch of HMMs with data y.unsqueeze(-2), :398-409 emissions with event_dim > 1): forward-backward, the Markov
    statistics and the emission update through HMM.update (models/HMM.py:120-152)."""
    from models.HMM import HMM
    g = torch.Generator().manual_seed(71)

    def switching(K, d, Tn, S, noise):
        A = torch.rand(K, K, generator=g) + 4 * torch.eye(K)
        A = A / A.sum(-1, keepdim=True)
        B = 2.0 * torch.randn(K, d, generator=g)
        z = torch.zeros(Tn, S, dtype=torch.long)
        z[0] = torch.randint(K, (S,), generator=g)
        for t in range(1, Tn):
            z[t] = torch.multinomial(A[z[t - 1]], 1, generator=g).squeeze(-1)
        return B[z] + noise * torch.randn(Tn, S, d, generator=g)

    cases = (("hmm_niw_k6", dict(K=4, d=2, Tn=30, S=20), (2,), (6,), lambda y: y),
             ("hmm_batch3_k6", dict(K=4, d=2, Tn=24, S=10), (2,), (3, 6), lambda y: y.unsqueeze(-2)),
             ("hmm_event32_k5", dict(K=5, d=6, Tn=20, S=15), (3, 2), (5,), lambda y: y.reshape(y.shape[:2] + (3, 2))))
    band = torch.tril(torch.ones(6, 6), 1) * torch.triu(torch.ones(6, 6), -1)      # a banded (left-right-ish) transition structure
    extra = {"hmm_masked_k6": dict(transition_mask=band), "hmm_ptemp2_k6": dict(ptemp=2.0),
             "hmm_masked_ptemp05_k6": dict(transition_mask=band, ptemp=0.5)}
    cases = cases + tuple((nm, dict(K=4, d=2, Tn=30, S=20), (2,), (6,), (lambda y: y)) for nm in extra)
    only = set(sys.argv[2:])
    for name, gen, ev, bs, shape in cases:
        if only and name not in only:
            switching(noise=0.3, **gen)          # keep the data generator's stream in step
            continue
        y = shape(switching(noise=0.3, **gen))
        torch.manual_seed(13)
        obs = dists.NormalInverseWishart(event_shape=ev, batch_shape=bs)
        kw = extra.get(name, {})
        m = HMM(obs, **kw)

        def st():
            s = niw_state(m.obs_dist, "obs.")
            s.update(dir_state(m.transition, "transition."))
            s.update(dir_state(m.initial, "initial."))
            return s
        out = {"y": T(y), "event_shape": np.asarray(ev), "batch_shape": np.asarray(bs),
               "ptemp": np.float64(kw.get("ptemp", 1.0))}
        if "transition_mask" in kw:
            out["transition_mask"] = T(kw["transition_mask"])
        out.update(tagged(st(), "init"))
        out["init/obs_logits"] = T(m.obs_logits(y))
        el = []
        for i in range(3):
            m.update(y, iters=1, lr=1.0)
            el.append(T(m.ELBO_last).astype(np.float64))
            if i == 0:
                out.update(tagged(st(), "iter1"))
                out["iter1/p"], out["iter1/logZ"], out["iter1/NA"] = T(m.p), T(m.logZ), T(m.NA)
        out.update(tagged(st(), "final"))
        out["final/p"], out["final/logZ"] = T(m.p), T(m.logZ)
        out["final/assignment"] = T(m.assignment()).astype(np.int32)
        out["ELBO"] = np.stack(el)
        save(name, **out)


def gen_arhmm_prxy():
    """models.ARHMM.ARHMM_prXY (models/ARHMM.py:35-46): the ARHMM driven by Gaussian beliefs about regressors and outputs."""
    from dists.MultivariateNormal_vector_format import MultivariateNormal_vector_format as MVN
    from models.ARHMM import ARHMM_prXY
    g = torch.Generator().manual_seed(53)
    K, n, p, Tn, S = 4, 2, 3, 40, 12
    Atrue = torch.randn(K, n, p, generator=g)
    btrue = torch.randn(K, n, generator=g)
    trans = torch.full((K, K), 0.1 / (K - 1)) + torch.eye(K) * (0.9 - 0.1 / (K - 1))
    z = torch.zeros(Tn, S, dtype=torch.long)
    z[0] = torch.randint(K, (S,), generator=g)
    for t in range(1, Tn):
        z[t] = torch.multinomial(trans[z[t - 1]], 1, generator=g).squeeze(-1)
    X = torch.randn(Tn, S, p, generator=g)
    Y = torch.einsum("tsij,tsj->tsi", Atrue[z], X) + btrue[z] + 0.2 * torch.randn(Tn, S, n, generator=g)
    Ax, Ay = 0.2 * torch.randn(Tn, S, 1, p, p, generator=g), 0.1 * torch.randn(Tn, S, 1, n, n, generator=g)
    Sx = Ax @ Ax.transpose(-1, -2) + 0.01 * torch.eye(p)
    Sy = Ay @ Ay.transpose(-1, -2) + 0.01 * torch.eye(n)
    mux, muy = X.unsqueeze(-2).unsqueeze(-1), Y.unsqueeze(-2).unsqueeze(-1)        # (T,S,1,p,1), (T,S,1,n,1)
    torch.manual_seed(16)
    m = ARHMM_prXY(K, n, p)
    out = {"mux": T(mux), "Sx": T(Sx), "muy": T(muy), "Sy": T(Sy), "K": K, "n": n, "p": p}

    def st():
        s = mnw_state(m.obs_dist, "obs.")
        s.update(dir_state(m.transition, "transition."))
        s.update(dir_state(m.initial, "initial."))
        return s
    out.update(tagged(st(), "init"))
    pX, pY = MVN(mu=mux, Sigma=Sx), MVN(mu=muy, Sigma=Sy)
    out["init/obs_logits"] = T(m.obs_logits((pX, pY)))
    el = []
    for i in range(3):
        m.update((pX, pY), iters=1, lr=1.0)
        el.append(float(m.ELBO_last))
        if i == 0:
            out.update(tagged(st(), "iter1"))
            out["iter1/p"], out["iter1/logZ"], out["iter1/NA"] = T(m.p), T(m.logZ), T(m.NA)
    out["ELBO"] = np.asarray(el, dtype=np.float64)
    save("arhmm_prxy_k4_n2_p3", **out)

def gen_arhmm_prxry():
    """models.ARHMM.ARHMM_prXRY (models/ARHMM.py:55-77): belief about latent regressors X stacked on observed regressors R,
    observed outputs Y (DynamicMarkovBlanketDiscovery's observation model)."""
    from dists.MultivariateNormal_vector_format import MultivariateNormal_vector_format as MVN
    from models.ARHMM import ARHMM_prXRY
    g = torch.Generator().manual_seed(57)
    K, n, p1, p2, Tn, S = 4, 2, 2, 1, 40, 12
    Atrue = torch.randn(K, n, p1 + p2, generator=g)
    trans = torch.full((K, K), 0.1 / (K - 1)) + torch.eye(K) * (0.9 - 0.1 / (K - 1))
    z = torch.zeros(Tn, S, dtype=torch.long)
    z[0] = torch.randint(K, (S,), generator=g)
    for t in range(1, Tn):
        z[t] = torch.multinomial(trans[z[t - 1]], 1, generator=g).squeeze(-1)
    XR = torch.randn(Tn, S, p1 + p2, generator=g)
    Y = torch.einsum("tsij,tsj->tsi", Atrue[z], XR) + 0.2 * torch.randn(Tn, S, n, generator=g)
    Ax = 0.2 * torch.randn(Tn, S, 1, p1, p1, generator=g)
    Sx = Ax @ Ax.transpose(-1, -2) + 0.01 * torch.eye(p1)
    mux = XR[..., :p1].unsqueeze(-2).unsqueeze(-1)                                  # (T,S,1,p1,1)
    R = XR[..., p1:].unsqueeze(-2).unsqueeze(-1)                                    # (T,S,1,p2,1)
    Yv = Y.unsqueeze(-2).unsqueeze(-1)                                              # (T,S,1,n,1)
    torch.manual_seed(17)
    m = ARHMM_prXRY(K, n, p1, p2)
    out = {"mux": T(mux), "Sx": T(Sx), "R": T(R), "Y": T(Yv), "K": K, "n": n, "p1": p1, "p2": p2}

    def st():
        s = mnw_state(m.obs_dist, "obs.")
        s.update(dir_state(m.transition, "transition."))
        s.update(dir_state(m.initial, "initial."))
        return s
    out.update(tagged(st(), "init"))
    XRY = (MVN(mu=mux, Sigma=Sx), R, Yv)
    out["init/obs_logits"] = T(m.obs_logits(XRY))
    el = []
    for i in range(3):
        m.update(XRY, iters=1, lr=1.0)
        el.append(float(m.ELBO_last))
        if i == 0:
            out.update(tagged(st(), "iter1"))
            out["iter1/p"], out["iter1/logZ"], out["iter1/NA"] = T(m.p), T(m.logZ), T(m.NA)
    out["ELBO"] = np.asarray(el, dtype=np.float64)
    save("arhmm_prxry_k4_n2_p21", **out)


# ---- diagonal-precision nodes: NormalGamma / GaussianMixtureModel(isotropic=True), MatrixNormalGamma /
# ---- MixtureofLinearTransforms(type='Gamma')

def gamma_state(g, pre):
    return {pre + k: T(getattr(g, k)) for k in ("alpha_0", "beta_0", "alpha", "beta")}


def ng_state(d, pre):
    out = {pre + k: T(getattr(d, k)) for k in ("lambda_mu_0", "lambda_mu", "mu_0", "mu")}
    out.update(gamma_state(d.gamma, pre + "gamma."))
    return out


def mng_state(d, pre):
    out = {pre + k: T(getattr(d, k)) for k in ("mu_0", "mu", "invV_0", "invV", "V", "logdetinvV", "logdetinvV_0")}
    out.update(gamma_state(d.invU.gamma, pre + "invU.gamma."))
    return out


def gen_diag():
    g = torch.Generator().manual_seed(51)
    # GaussianMixtureModel(isotropic=True): NormalGamma components (models/GaussianMixtureModel.py:8-11)
    for name, d, K, N, iters, seed in (("gmm_iso_d8_k6", 8, 6, 600, 8, 4), ("gmm_iso_d64_k32", 64, 32, 1024, 4, 5)):
        mu = (2.0 if d == 8 else 0.5) * torch.randn(K, d, generator=g)
        sd = 0.5 + torch.rand(K, d, generator=g)
        z = torch.randint(K, (N,), generator=g)
        X = mu[z] + sd[z] * torch.randn(N, d, generator=g)
        torch.manual_seed(seed)
        m = models.GaussianMixtureModel(K, d, isotropic=True)
        m.initialize(X)

        def st():
            s = ng_state(m.dist, "dist.")
            s.update(dir_state(m.pi, "pi."))
            return s
        out = {"X": T(X), "nc": K, "iters": iters, "lr": 1.0}
        out.update(tagged(st(), "init"))
        out["init/Elog_like"] = T(m.dist.Elog_like(X.unsqueeze(-2)))
        out["init/KL"] = T(m.KLqprior())
        el = []
        for i in range(iters):
            m.update(X, 1, 1.0)
            el.append(float(m.ELBO_last))
            if i == 0:
                out.update(tagged(st(), "iter1"))
                out["iter1/logZ"], out["iter1/NA"], out["iter1/KL"], out["iter1/p"] = T(m.logZ), T(m.NA), T(m.KLqprior()), T(m.p)
        out.update(tagged(st(), "final"))
        out["final/assignment"] = T(m.assignment()).astype(np.int32)
        out["final/Elog_like"] = T(m.Elog_like(X))
        out["final/KL"] = T(m.KLqprior())
        out["ELBO"] = np.asarray(el, dtype=np.float64)
        save(name, **out)
    # NormalGamma.raw_update with beta forgetting and lr < 1, then the unit-weight branch (dists/NormalGamma.py:41-73)
    torch.manual_seed(6)
    d = dists.NormalGamma((3,), (4,), scale=0.7)
    out = tagged(ng_state(d, ""), "init")
    for i in range(3):
        X = torch.randn(50, 1, 3, generator=g) * 1.3 + 0.4
        pw = torch.softmax(torch.randn(50, 4, generator=g), -1)
        d.raw_update(X, pw, lr=0.6, beta=0.9)
        out[f"X{i}"], out[f"p{i}"] = T(X), T(pw)
        out.update(tagged(ng_state(d, ""), f"step{i}"))
        out[f"step{i}/SExx"], out[f"step{i}/SEx"], out[f"step{i}/N"] = T(d.SExx), T(d.SEx), T(d.N)
    Xb = torch.randn(40, 4, 3, generator=g)
    d.raw_update(Xb, None, lr=1.0, beta=None)
    out["Xb"] = T(Xb)
    out.update(tagged(ng_state(d, ""), "pnone"))
    out["final/KL"] = T(d.KLqprior())
    out["final/Elog_like"] = T(d.Elog_like(torch.as_tensor(out["X2"])))
    save("ng_beta_lr", **out)
    # MatrixNormalGamma steps (transforms/MatrixNormalGamma.py:87-243)
    for pad in (True, False):
        torch.manual_seed(9)
        n, p, K, N = 4, 5, 3, 300
        d = transforms.MatrixNormalGamma(event_shape=(n, p), batch_shape=(K,), scale=0.8, pad_X=pad)
        out = {"n": n, "p": p, "K": K, "pad_X": int(pad)}
        out.update(tagged(mng_state(d, ""), "init"))
        X = torch.randn(N, 1, p, 1, generator=g)
        Wt = torch.randn(K, n, p, generator=g)
        z = torch.randint(K, (N,), generator=g)
        Y = (Wt[z] @ X.squeeze(1)).unsqueeze(1) + 0.3 * torch.randn(N, 1, n, 1, generator=g) + 0.5
        r = torch.softmax(torch.randn(N, K, generator=g), -1)
        out["X"], out["Y"], out["r"] = T(X), T(Y), T(r)
        out["init/Elog_like"] = T(d.Elog_like(X, Y))
        out["init/KL"] = T(d.KLqprior())
        d.raw_update(X, Y, p=r, lr=1.0, beta=None)
        out.update(tagged(mng_state(d, ""), "step0"))
        out["step0/Elog_like"] = T(d.Elog_like(X, Y))
        out["step0/KL"] = T(d.KLqprior())
        d.raw_update(X, Y, p=r, lr=0.5, beta=0.8)
        out.update(tagged(mng_state(d, ""), "step1"))
        out["step1/SExx"], out["step1/SEyx"], out["step1/SEyy"], out["step1/N"] = T(d.SExx), T(d.SEyx), T(d.SEyy), T(d.N)
        out["step1/Elog_like"] = T(d.Elog_like(X, Y))
        out["step1/KL"] = T(d.KLqprior())
        save(f"mng_n4_p5_k3_pad{int(pad)}", **out)
    # MixtureofLinearTransforms(type='Gamma') trajectories
    for name, n, p, K, N, iters in (("molt_gamma_n3_p4_k5", 3, 4, 5, 600, 5), ("molt_gamma_n32_p32_k8", 32, 32, 8, 768, 3)):
        torch.manual_seed(10)
        m = transforms.MixtureofLinearTransforms(n, p, K, pad_X=True, type='Gamma')
        X = torch.randn(N, p, generator=g)
        Wt = torch.randn(K, n, p, generator=g) / np.sqrt(p)
        b = torch.randn(K, n, generator=g)
        z = torch.randint(K, (N,), generator=g)
        Y = torch.einsum("nij,nj->ni", Wt[z], X) + b[z] + 0.1 * torch.randn(N, n, generator=g)

        def st():
            s = mng_state(m.W, "W.")
            s.update(dir_state(m.pi, "pi."))
            return s
        out = {"X": T(X), "Y": T(Y), "n": n, "p": p, "K": K, "iters": iters}
        out.update(tagged(st(), "init"))
        el = []
        for i in range(iters):
            m.raw_update(X.unsqueeze(-1), Y.unsqueeze(-1), iters=1, lr=1.0)
            el.append(float(m.ELBO_last))
            if i == 0:
                out.update(tagged(st(), "iter1"))
                out["iter1/p"], out["iter1/logZ"] = T(m.p), T(m.logZ)
        out.update(tagged(st(), "final"))
        out["final/p"], out["final/logZ"] = T(m.p), T(m.logZ)
        out["final/assignment"] = T(m.assignment()).astype(np.int32)
        out["final/KL"] = T(m.KLqprior())
        out["ELBO"] = np.asarray(el, dtype=np.float64)
        save(name, **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "round2":    # only the fixtures added in round 2
        gen_arhmm_prxry()
        gen_diag()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "hmm":       # the HMM layouts added late in round 2
        gen_hmm_variants()
        sys.exit(0)
    gen_gmm()
    gen_niw_variants()
    gen_mnw()
    gen_molt()
    gen_molt_predict()
    gen_molt_given()
    gen_arhmm()
    gen_hmm_variants()
    gen_arhmm_prxy()
    gen_diag()
    gen_arhmm_prxry()
