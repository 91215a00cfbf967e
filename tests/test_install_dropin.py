"""install() must not take anything away from a pyVBMP tree (SURVEY.md §8b, Appendix D "untouched models still run
through the boundary").  These tests need the reference itself: they run where it is importable
(``PYVBMP_REFERENCE``, /root/reference in the build container, or the offline install under baseline/_ref) and skip
elsewhere.

CPU part (no GPU needed): after install() the reference's out-of-scope models — LinearDynamicalSystems,
DynamicMarkovBlanketDiscovery, dMixtureofLinearTransforms, masked MatrixNormalWishart, the message-passing methods —
construct and update, and give the SAME numbers as before install() from the same seed (CPU-resident nodes keep the
reference's own code; nothing in this library computes on the CPU).

GPU part (``-m gpu``): the reference's OWN GaussianMixtureModel / MixtureofLinearTransforms / ARHMM classes, constructed
under ``torch.set_default_device('cuda')`` after install(), run on libvbmp_b200.so and reproduce the golden fixtures
made by the unmodified reference on the CPU.
"""
import os
import sys

import numpy as np
import pytest
import torch

import pyvbmp_b200 as V
from _util import load_golden, tag, assert_close, assert_maxabs

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# an importable, unmodified pyVBMP tree: $PYVBMP_REFERENCE, the build container's /root/reference, or the offline install
# under baseline/_ref (git-ignored; `pip install --no-deps --target baseline/_ref <reference>`) which travels
# to the GPU box
REF = next((p for p in (os.environ.get("PYVBMP_REFERENCE"), "/root/reference", os.path.join(_ROOT, "baseline", "_ref"))
            if p and os.path.isdir(os.path.join(p, "dists")) and os.path.isdir(os.path.join(p, "models"))), "")
HAVE_REF = bool(REF)
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="the pyVBMP reference tree is not available here")


@pytest.fixture
def ref_tree():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import dists, transforms, models          # noqa: F401,E401
    yield sys.modules["dists"], sys.modules["transforms"], sys.modules["models"]
    V.uninstall()


def _lds_run(models):
    torch.manual_seed(3)
    T, B, obs, hid, cd, rd = 12, 5, 4, 2, 2, 2
    y, u, r = torch.randn(T, B, obs), torch.randn(T, B, cd), torch.randn(T, B, rd)
    lds = models.LinearDynamicalSystems((obs,), hid, cd, rd, latent_noise='shared')
    lds.update(y, u, r, iters=2, lr=1.0, verbose=False)
    return lds.px.mean().clone(), lds.ELBO().clone() if hasattr(lds, "ELBO") else torch.zeros(())


def _dmolt_run(transforms):
    torch.manual_seed(4)
    n, p, nc, N = 3, 4, 3, 200
    X, Y = torch.randn(N, p), torch.randn(N, n)
    m = transforms.dMixtureofLinearTransforms(n, p, nc, batch_shape=(), pad_X=True)
    m.raw_update(X, Y, iters=2, lr=1.0, verbose=False)
    pY, pr = m.predict(X)
    return pY.mean().clone(), pr.clone()


def _dmbd_run(models):
    torch.manual_seed(5)
    T, B, nobj, od = 8, 3, 4, 2
    data = torch.randn(T, B, nobj, od)
    m = models.DynamicMarkovBlanketDiscovery(obs_shape=(nobj, od), role_dims=(2, 2, 2), hidden_dims=(2, 2, 2),
                                             batch_shape=(), number_of_objects=1)
    m.update(data, None, None, iters=1, latent_iters=1, lr=0.5, verbose=False)
    return m.px.mean().clone(), m.obs_model.p.clone()


def _masked_mnw_run(transforms):
    torch.manual_seed(6)
    n, p, K, N = 3, 4, 2, 50
    mask = torch.ones(n, p, dtype=torch.bool)
    mask[0, 1] = False
    w = transforms.MatrixNormalWishart(event_shape=(n, p), batch_shape=(K,), mask=mask)
    X, Y = torch.randn(N, 1, p, 1), torch.randn(N, 1, n, 1)
    w.raw_update(X, Y, p=torch.rand(N, K))
    return w.mu.clone(), w.Elog_like(X, Y).clone()


def _message_passing_run(transforms, dists):
    torch.manual_seed(7)
    n, p, K, N = 3, 4, 2, 20
    w = transforms.MatrixNormalWishart(event_shape=(n, p), batch_shape=(K,), pad_X=True)
    X, Y = torch.randn(N, 1, p, 1), torch.randn(N, 1, n, 1)
    w.raw_update(X, Y, p=torch.rand(N, K))
    invS, invSmu, Res = w.Elog_like_X(Y)
    pX = dists.MultivariateNormal_vector_format(mu=X, Sigma=torch.eye(p).expand(N, 1, p, p) * 0.1)
    out = w.forward(pX)
    pY = out[0] if isinstance(out, tuple) else out
    return invS.clone(), invSmu.clone(), Res.clone(), pY.mean().clone()


@needs_ref
@pytest.mark.parametrize("which", ["lds", "dmolt", "dmbd", "masked_mnw", "message_passing"])
def test_install_keeps_out_of_scope_models_working_on_cpu(ref_tree, which):
    dists, transforms, models = ref_tree
    run = {"lds": lambda: _lds_run(models), "dmolt": lambda: _dmolt_run(transforms), "dmbd": lambda: _dmbd_run(models),
           "masked_mnw": lambda: _masked_mnw_run(transforms),
           "message_passing": lambda: _message_passing_run(transforms, dists)}[which]
    V.uninstall()
    before = run()
    n = V.install(REF)
    assert n > 0
    cls = V.installed_classes()
    assert transforms.MatrixNormalWishart is cls["MatrixNormalWishart"]
    assert issubclass(cls["MatrixNormalWishart"], V.MatrixNormalWishart)
    after = run()
    for a, b in zip(before, after):
        assert a.shape == b.shape
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), (which, float((a - b).abs().max()))
    V.uninstall()
    assert transforms.MatrixNormalWishart is not cls["MatrixNormalWishart"]


@needs_ref
def test_installed_classes_are_subclasses_of_both(ref_tree):
    dists, transforms, models = ref_tree
    V.uninstall()
    ref_niw, ref_w, ref_mnw = dists.NormalInverseWishart, dists.Wishart, transforms.MatrixNormalWishart
    V.install(REF)
    cls = V.installed_classes()
    for key, ref, ours in (("NormalInverseWishart", ref_niw, V.NormalInverseWishart), ("Wishart", ref_w, V.Wishart),
                           ("MatrixNormalWishart", ref_mnw, V.MatrixNormalWishart)):
        assert issubclass(cls[key], ref) and issubclass(cls[key], ours)
    # the reference's model modules now construct installed nodes
    torch.manual_seed(0)
    g = models.GaussianMixtureModel(4, 3)
    assert isinstance(g.dist, cls["NormalInverseWishart"]) and isinstance(g.dist.invU, cls["Wishart"])
    h = models.ARHMM(3, 2, 2)
    assert isinstance(h.obs_dist, cls["MatrixNormalWishart"]) and isinstance(h.obs_dist.invU, cls["Wishart"])
    # masked nodes are plain reference objects, message-passing methods are the reference's
    w = transforms.MatrixNormalWishart(event_shape=(2, 2), batch_shape=(), mask=torch.ones(2, 2, dtype=torch.bool))
    assert type(w) is ref_mnw
    assert cls["MatrixNormalWishart"].forward is ref_mnw.forward
    assert not hasattr(V.MatrixNormalWishart((2, 2), (3,)), "forward")
    # identical RNG consumption: same seed -> same initial state as the reference class
    V.uninstall()
    torch.manual_seed(11)
    a = models.GaussianMixtureModel(5, 3)
    V.install(REF)
    torch.manual_seed(11)
    b = models.GaussianMixtureModel(5, 3)
    assert torch.equal(a.dist.mu, b.dist.mu) and torch.equal(a.pi.alpha, b.pi.alpha)
    assert torch.allclose(a.dist.invU.U, b.dist.invU.U) and torch.allclose(a.dist.invU.logdet_invU, b.dist.invU.logdet_invU)


# -------------------------------------------------------------------------------------------------------------------
# GPU: the reference's own model classes on the CUDA path, against the fixtures the unmodified reference produced
# -------------------------------------------------------------------------------------------------------------------

def _set_state(obj, flat, device):
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


def _purge_reference_modules():
    for name in [n for n in sys.modules if n in ("dists", "transforms", "models", "utils")
                 or n.startswith(("dists.", "transforms.", "models.", "utils."))]:
        del sys.modules[name]


@pytest.fixture
def cuda_default():
    """The reference creates constants — and the tensors in its mutable default arguments (dists/Dirichlet.py:4,
    dists/NormalInverseWishart.py:7-10) — on the DEFAULT device at import time, so a GPU run must import it under
    torch.set_default_device('cuda'): purge, set the device, import, install."""
    V.uninstall()
    _purge_reference_modules()
    old = torch.empty(0).device
    torch.set_default_device("cuda:0")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import dists, transforms, models          # noqa: F401,E401
    V.install(REF)
    yield sys.modules["dists"], sys.modules["transforms"], sys.modules["models"]
    torch.set_default_device(old)
    V.uninstall()
    _purge_reference_modules()


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["gmm_d2_k6", "gmm_d64_k32_overlap"])
def test_reference_gmm_runs_on_the_cuda_path(cuda_default, name):
    dists, transforms, models = cuda_default
    from pyvbmp_b200 import _lib
    fix = load_golden(name)
    X = torch.as_tensor(fix["X"]).to("cuda:0")
    nc, iters, lr = int(fix["nc"]), int(fix["iters"]), float(fix["lr"])
    torch.manual_seed(0)
    m = models.GaussianMixtureModel(nc, X.shape[-1])
    assert type(m).__module__.startswith("models.") and m.dist.mu.is_cuda
    _set_state(m, tag(fix, "init"), "cuda:0")
    n0 = _lib.LAUNCHES
    elbo = []
    for i in range(iters):
        m.update(X, 1, lr)                                    # the reference's Mixture.update loop (dists/Mixture.py:54-62)
        elbo.append(float(m.ELBO_last))
        if i == 0:
            it1 = tag(fix, "iter1")
            for k in ("dist.mu", "dist.lambda_mu", "dist.invU.invU", "dist.invU.U", "dist.invU.nu", "pi.alpha"):
                assert_close(_get(m, k), it1[k], 1e-4, k)
            assert_close(m.NA, it1["NA"], 1e-4, "NA")
    assert _lib.LAUNCHES - n0 >= 5 * iters                    # K1, K2, KL, K3, K5 every iteration: the kernels ran
    ref = fix["ELBO"]
    assert np.max(np.abs(np.array(elbo) - ref) / np.abs(ref)) < 1e-4, (elbo, ref)
    assert (m.assignment().cpu().numpy() == fix["final/assignment"]).mean() > 0.999


@needs_ref
@pytest.mark.gpu
def test_reference_molt_and_arhmm_run_on_the_cuda_path(cuda_default):
    dists, transforms, models = cuda_default
    from pyvbmp_b200 import _lib
    fix = load_golden("molt_n32_p32_k8")
    n, p, K, iters = (int(fix[k]) for k in ("n", "p", "K", "iters"))
    torch.manual_seed(0)
    m = transforms.MixtureofLinearTransforms(n, p, K)
    _set_state(m, tag(fix, "init"), "cuda:0")
    X, Y = torch.as_tensor(fix["X"]).unsqueeze(-1).to("cuda:0"), torch.as_tensor(fix["Y"]).unsqueeze(-1).to("cuda:0")
    n0 = _lib.LAUNCHES
    elbo = []
    for _ in range(iters):
        m.raw_update(X, Y, iters=1)
        elbo.append(float(m.ELBO_last))
    assert _lib.LAUNCHES - n0 >= 5 * iters
    assert np.max(np.abs(np.array(elbo) - fix["ELBO"]) / np.abs(fix["ELBO"])) < 1e-4
    assert (m.assignment().cpu().numpy() == fix["final/assignment"]).mean() > 0.995
    # the message-passing side of the same (CUDA-resident, updated) node is still the reference's code
    pY, pr = m.predict(X)
    assert pr.shape == (X.shape[0], K) and torch.isfinite(pY.mean()).all()

    fix = load_golden("arhmm_k4_n2_p3")
    K, n, p = int(fix["K"]), int(fix["n"]), int(fix["p"])
    torch.manual_seed(0)
    h = models.ARHMM(K, n, p)
    _set_state(h, {k.replace("obs.", "obs_dist."): v for k, v in tag(fix, "init").items()}, "cuda:0")
    X, Y = torch.as_tensor(fix["X"]).to("cuda:0"), torch.as_tensor(fix["Y"]).to("cuda:0")
    elbo = []
    n0 = _lib.LAUNCHES
    for i in range(4):
        h.update((X, Y), iters=1)
        elbo.append(float(h.ELBO_last))
        if i == 0:
            it1 = tag(fix, "iter1")
            assert_maxabs(h.p.cpu(), it1["p"], 2e-4, "p")
            for k in ("obs.mu", "obs.invV", "obs.invU.invU", "transition.alpha", "initial.alpha"):
                assert_close(_get(h, k.replace("obs.", "obs_dist.")), it1[k], 1e-4, k)
    assert _lib.LAUNCHES - n0 >= 4 * 5
    assert np.max(np.abs(np.array(elbo) - fix["ELBO"]) / np.abs(fix["ELBO"])) < 1e-4


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["hmm_niw_k6", "hmm_batch3_k6", "hmm_event32_k5", "hmm_masked_k6", "hmm_ptemp2_k6",
                                  "hmm_masked_ptemp05_k6"])
def test_reference_hmm_runs_on_the_cuda_path(cuda_default, name):
    """The reference's own models.HMM over an installed NormalInverseWishart node — plain, a batch of HMMs
    (tests/test_models.py:353-356) and emissions with event_dim > 1 (:398-409) — against the unmodified reference's outputs."""
    dists, transforms, models = cuda_default
    from pyvbmp_b200 import _lib
    fix = load_golden(name)
    ev, bs = tuple(int(v) for v in fix["event_shape"]), tuple(int(v) for v in fix["batch_shape"])
    torch.manual_seed(0)
    kw = {"ptemp": float(fix["ptemp"])}
    if "transition_mask" in fix:
        kw["transition_mask"] = torch.as_tensor(fix["transition_mask"]).to("cuda:0")
    h = models.HMM(dists.NormalInverseWishart(event_shape=ev, batch_shape=bs), **kw)
    assert type(h).__module__.startswith("models.") and h.obs_dist.mu.is_cuda
    _set_state(h, {k.replace("obs.", "obs_dist."): v for k, v in tag(fix, "init").items()}, "cuda:0")
    y = torch.as_tensor(fix["y"]).to("cuda:0")
    n0 = _lib.LAUNCHES
    elbo = []
    for i in range(3):
        h.update(y, iters=1)
        elbo.append(h.ELBO_last.detach().cpu().double().numpy())
        if i == 0:
            it1 = tag(fix, "iter1")
            assert float((h.p.cpu() - it1["p"].cpu()).abs().max()) < 2e-4
            for k in ("obs.mu", "obs.lambda_mu", "obs.invU.invU", "obs.invU.nu", "transition.alpha", "initial.alpha"):
                assert_close(_get(h, k.replace("obs.", "obs_dist.")), it1[k], 1e-4, k)
            assert_close(h.NA, it1["NA"], 1e-4, "NA")
    assert _lib.LAUNCHES - n0 >= 4 * 3                        # K1, K2, K6, K3 / K5 every iteration
    assert np.max(np.abs(np.stack(elbo) - fix["ELBO"]) / np.abs(fix["ELBO"])) < 1e-4
    assert (h.assignment().cpu().numpy() == fix["final/assignment"]).mean() > 0.995


# -------------------------------------------------------------------------------------------------------------------
# GPU: the reference's OTHER models — hierarchical / tensor HMMs, mixtures with replica and extra event dims, LDS,
# dMixtureofLinearTransforms — drive the installed nodes with layouts the three target models never produce (SURVEY.md
# Appendix D last row, Appendix E).  Three runs of the same scenario from the same initial state:
#   (A) the unmodified reference on the CPU                      -> the truth
#   (B) the unmodified reference on CUDA, nothing installed      -> can the reference itself run this on a GPU at all?
#   (C) the reference on CUDA with install()                     -> the drop-in
# If (B) runs, (C) must run and agree with (A); if the reference's own code is not GPU-clean for a scenario, it is skipped.
# -------------------------------------------------------------------------------------------------------------------

def _walk_tensors(obj, fn, seen=None, depth=0):
    """Apply fn(owner, key, tensor) to every tensor attribute of a pyVBMP object graph (objects, lists, tuples, dicts)."""
    seen = set() if seen is None else seen
    if id(obj) in seen or depth > 8:
        return
    seen.add(id(obj))
    if isinstance(obj, (list, tuple)):
        items = list(enumerate(obj))
    elif isinstance(obj, dict):
        items = list(obj.items())
    elif hasattr(obj, "__dict__") and type(obj).__module__.split(".")[0] in ("dists", "transforms", "models", "pyvbmp_b200"):
        items = list(vars(obj).items())
    else:
        return
    for k, v in items:
        if isinstance(v, torch.Tensor):
            fn(obj, k, v)
        else:
            _walk_tensors(v, fn, seen, depth + 1)


def _snapshot(model):
    out = []
    _walk_tensors(model, lambda o, k, t: out.append(t.detach().cpu().clone()))
    return out


def _restore(model, snap, device):
    it = iter(snap)

    def put(o, k, t):
        v = next(it)
        assert tuple(v.shape) == tuple(t.shape), (type(o).__name__, k, v.shape, t.shape)
        if isinstance(o, list):
            o[k] = v.to(device)
        elif isinstance(o, dict):
            o[k] = v.to(device)
        elif not isinstance(o, tuple):
            setattr(o, k, v.to(device))
    _walk_tensors(model, put)


def _switching(g, K, d, T, S):
    A = torch.rand(K, K, generator=g) + 4 * torch.eye(K)
    A = A / A.sum(-1, keepdim=True)
    B = 2.0 * torch.randn(K, d, generator=g)
    z = torch.zeros(T, S, dtype=torch.long)
    z[0] = torch.randint(K, (S,), generator=g)
    for t in range(1, T):
        z[t] = torch.multinomial(A[z[t - 1]], 1, generator=g).squeeze(-1)
    return B[z] + 0.3 * torch.randn(T, S, d, generator=g)


def _scenario(which, mods, data, dev):
    """Build the model of scenario `which` on the current default device; returns (model, step, outputs)."""
    dists, transforms, models = mods
    if which == "hhmm":                      # tests/test_models.py:320-326
        from models.HHMM import HHMM
        m = HHMM(dists.NormalInverseWishart(event_shape=(2,), batch_shape=(2, 3, 2)), 3)
        return m, (lambda: m.update(data["y"].to(dev), iters=1, lr=1)), (lambda: (m.p, m.ELBO_last))
    if which == "tensor_hmm":                # tests/test_models.py:340-347
        from models.Tensor_HMM import Tensor_HMM
        m = Tensor_HMM(dists.NormalInverseWishart(event_shape=(2,), batch_shape=(2, 3, 2)), event_shape=(2, 3, 2))
        return m, (lambda: m.update(data["y"].to(dev), iters=1, lr=1)), (lambda: (m.p, m.ELBO_last))
    if which == "mixture_joint":             # tests/test_dists.py:284-288: softmax jointly over (G, K), event_dim > 1 emissions
        m = dists.Mixture(dists.NormalInverseWishart(event_shape=(3, 2), batch_shape=(2, 4)), event_shape=(2, 4))
        return m, (lambda: m.update(data["x32"].to(dev), iters=1, lr=1)), (lambda: (m.p, m.ELBO_last, m.NA))
    if which == "mixture_replicas":          # tests/test_dists.py:261-276: G independent mixtures, each with its own column
        m = dists.Mixture(dists.NormalInverseWishart(event_shape=(2,), batch_shape=(3, 4)), event_shape=(4,))
        return m, (lambda: m.update(data["xg"].to(dev), iters=1, lr=1)), (lambda: (m.p, m.ELBO_last, m.NA))
    if which == "lds":                       # models/LinearDynamicalSystems.py:96,152-153 call MatrixNormalWishart.ss_update
        m = models.LinearDynamicalSystems((4,), 2, 2, 2, latent_noise='shared')
        y, u, r = data["ly"].to(dev), data["lu"].to(dev), data["lr"].to(dev)
        return m, (lambda: m.update(y, u, r, iters=1, lr=1.0, verbose=False)), (lambda: (m.px.mean(), m.logZ))
    if which == "dmolt":                     # transforms/dMixtureofLinearTransforms.py:42,54
        m = transforms.dMixtureofLinearTransforms(3, 4, 3, batch_shape=(), pad_X=True)
        X, Y = data["dx"].to(dev), data["dy"].to(dev)
        return m, (lambda: m.raw_update(X, Y, iters=1, lr=1.0, verbose=False)), (lambda: (m.predict(X)[1],))
    raise KeyError(which)


def _import_reference(device):
    V.uninstall()
    _purge_reference_modules()
    torch.set_default_device(device)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import dists, transforms, models          # noqa: F401,E401
    return sys.modules["dists"], sys.modules["transforms"], sys.modules["models"]


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("which", ["hhmm", "tensor_hmm", "mixture_joint", "mixture_replicas", "lds", "dmolt"])
def test_other_reference_models_on_cuda_match_the_cpu_reference(which):
    from pyvbmp_b200 import _lib
    old = torch.empty(0).device
    g = torch.Generator().manual_seed(91)
    data = {"y": _switching(g, 5, 2, 16, 6), "x32": torch.randn(300, 3, 2, generator=g) * 1.5,
            "xg": torch.randn(300, 3, 2, generator=g) * 1.5 + torch.randint(3, (300, 1, 1), generator=g) * 2.0,
            "ly": torch.randn(12, 5, 4, generator=g), "lu": torch.randn(12, 5, 2, generator=g),
            "lr": torch.randn(12, 5, 2, generator=g), "dx": torch.randn(200, 4, generator=g),
            "dy": torch.randn(200, 3, generator=g)}
    iters = 2
    try:
        # (A) the unmodified reference on the CPU
        mods = _import_reference("cpu")
        torch.manual_seed(17)
        m, step, outs = _scenario(which, mods, data, "cpu")
        snap = _snapshot(m)
        for _ in range(iters):
            step()
        truth = [torch.as_tensor(o).detach().cpu().double().clone() for o in outs()]

        def on_cuda(install):
            mods = _import_reference("cuda:0")
            if install:
                assert V.install(REF) > 0
            torch.manual_seed(17)
            m, step, outs = _scenario(which, mods, data, "cuda:0")
            _restore(m, snap, "cuda:0")
            for _ in range(iters):
                step()
            return [torch.as_tensor(o).detach().cpu().double().clone() for o in outs()]
        # (B) the unmodified reference on CUDA: is the reference's own code GPU-clean for this scenario?
        try:
            plain = on_cuda(False)
        except Exception as e:                 # noqa: BLE001
            pytest.skip(f"the unmodified reference does not run this scenario on CUDA itself: {type(e).__name__}: {e}"[:200])
        # (C) the drop-in
        n0 = _lib.LAUNCHES
        ours = on_cuda(True)
        launched = _lib.LAUNCHES - n0
    finally:
        torch.set_default_device(old)
        V.uninstall()
        _purge_reference_modules()
    for a, b, c in zip(truth, plain, ours):
        assert a.shape == c.shape
        scale = max(float(a.abs().max()), 1e-30)
        err_ours, err_plain = float((c - a).abs().max()) / scale, float((b - a).abs().max()) / scale
        # the drop-in may not be further from the CPU reference than 1e-4 (or than the reference's own CUDA run, if that
        # is already further: a free-running fp32 trajectory on another device)
        assert err_ours <= max(2e-4, 3 * err_plain), (which, err_ours, err_plain)
    if which != "dmolt":
        assert launched > 0, "install() was active but no kernel of the library ran"
