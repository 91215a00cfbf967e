"""GPU kernel-level tests through the C ABI: the tcgen05 (UMMA) E-step / Gram kernels and the CUDA-core kernels
against an fp64 torch restatement of the same op on the same seeded inputs (tolerances: BASELINE.json's 1e-4
relative on statistics / log-normalisers; logits to fp32 round-off of their magnitude; argmax identical)."""
import pytest
import torch

from pyvbmp_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _problem(N, d0, d1, K, seed=0, spread=1.0, fscale=None):
    """fscale: per-feature units (D,) applied to the data and, consistently, to the component parameters."""
    dev = torch.device(DEV)
    g = torch.Generator(device=dev).manual_seed(seed)
    D = d0 + d1
    Dp = _lib.pad_dim(D)
    mu = spread * torch.randn(K, D, generator=g, device=dev)
    z = mu[torch.randint(K, (N,), generator=g, device=dev)] + torch.randn(N, D, generator=g, device=dev)
    A = torch.randn(K, D, D, generator=g, device=dev) / D ** 0.5
    invU = A @ A.transpose(-1, -2) + 0.5 * torch.eye(D, device=dev)
    if fscale is not None:
        fscale = fscale.to(dev)
        mu = mu * fscale
        z = z * fscale
        invU = invU * fscale[:, None] * fscale[None, :]
    nu = D + 2 + 10 * torch.rand(K, generator=g, device=dev)
    lam = 1 + torch.rand(K, generator=g, device=dev)
    lp = torch.log_softmax(torch.randn(K, generator=g, device=dev), 0)
    W, m, cst, info = _lib.niw_prep(invU.contiguous(), mu.contiguous(), nu, lam, lp, K, D, Dp)
    assert int(info.abs().max()) == 0
    z0 = z[:, :d0].contiguous()
    z1 = z[:, d0:].contiguous() if d1 else None
    y = torch.einsum("ni,kij->nkj", z.double(), W[:, :D, :].double()) - m.double()[None]
    L = cst.double()[None] - 0.5 * (y * y).sum(-1)
    return z, z0, z1, W, m, cst, Dp, L


CASES = [(5000, 64, 0, 256), (66000, 64, 0, 256), (3001, 32, 32, 64), (4099, 16, 16, 32), (2500, 16, 0, 8),
         (6000, 48, 0, 20), (300, 64, 0, 16),
         # edges of the tensor-core kernels' shape windows: one row past a tile / chunk, smallest and largest K, ragged
         # feature counts that need zero padding, K not a multiple of the 128-component block
         (257, 64, 0, 4), (2048, 8, 8, 4), (2049, 64, 0, 512), (4097, 20, 12, 36), (9000, 64, 0, 132), (2304, 32, 0, 260),
         # 64 < D <= 128 (padded to 128): tcgen05 kernels with fp16 operands — two accumulator buffers in the E-step, 8385
         # pair columns in 44 blocks and a shallower chunk ring in the Gram
         (3000, 96, 0, 64), (4500, 128, 0, 256), (2500, 64, 64, 64), (2100, 72, 40, 12),
         # K % 4 != 0: the tensor-core kernels run on K padded by inert components (api.cu, kq_of)
         (5000, 64, 0, 50), (2600, 32, 32, 10), (3000, 128, 0, 30), (4096, 16, 0, 7), (2300, 64, 0, 257),
         # K > 512: the E-step runs per block of 512 components (logits with row stride K) + vbmp_softmax_rows
         (3000, 64, 0, 640), (12000, 32, 0, 1000), (2200, 16, 16, 770)]


@pytest.mark.parametrize("N,d0,d1,K", CASES)
@pytest.mark.parametrize("simt", [0, 1])
def test_estep_kernels(N, d0, d1, K, simt):
    z, z0, z1, W, m, cst, Dp, L = _problem(N, d0, d1, K)
    xg = torch.zeros(1, dtype=torch.int32, device=DEV)
    lz = torch.logsumexp(L, -1)
    P = (L - lz[:, None]).exp()
    old = _lib.FORCE_SIMT
    _lib.FORCE_SIMT = simt
    try:
        lg = _lib.estep(z0, z1, N, 1, xg, W, m, cst, 1, K, Dp, 0).view(N, K)
        p, lzn, NA, lZ = _lib.estep(z0, z1, N, 1, xg, W, m, cst, 1, K, Dp, 1)
    finally:
        _lib.FORCE_SIMT = old
    p = p.view(N, K)
    scale = float(L.abs().max())
    assert float((lg.double() - L).abs().max()) <= 2e-6 * scale          # fp32 round-off of the logit magnitude
    assert float((lzn.view(N).double() - lz).abs().max()) <= 2e-6 * scale
    assert float((p.double() - P).abs().max()) <= 2e-3
    assert float(((NA.view(K).double() - P.sum(0)).abs() / P.sum(0).clamp_min(1.0)).max()) <= 1e-4
    assert abs(float(lZ.double().sum() - lz.sum())) <= 1e-5 * abs(float(lz.sum()))
    # argmax: identical wherever the fp64 top-2 margin exceeds the fp32 logit noise
    top2 = L.topk(2, -1).values
    safe = (top2[:, 0] - top2[:, 1]) > 4e-6 * scale
    assert bool((p.argmax(-1) == P.argmax(-1))[safe].all())
    assert int((p.argmax(-1) != P.argmax(-1)).sum()) <= max(1, N // 2000)


@pytest.mark.parametrize("N,d0,d1,K", CASES)
@pytest.mark.parametrize("simt", [0, 1])
def test_gram_kernels(N, d0, d1, K, simt):
    z, z0, z1, W, m, cst, Dp, L = _problem(N, d0, d1, K, seed=1)
    D = d0 + d1
    xg = torch.zeros(1, dtype=torch.int32, device=DEV)
    P = (L - torch.logsumexp(L, -1)[:, None]).exp()
    Pf = P.float().contiguous()
    zt = torch.cat([z.double(), torch.ones(N, 1, device=DEV, dtype=torch.float64)], 1)
    Gref = torch.zeros(K, D + 1, D + 1, device=DEV, dtype=torch.float64)
    for a in range(0, N, 4096):
        Gref += torch.einsum("nk,ni,nj->kij", Pf[a:a + 4096].double(), zt[a:a + 4096], zt[a:a + 4096])
    old = _lib.FORCE_SIMT
    _lib.FORCE_SIMT = simt
    try:
        G = _lib.gram(z0, z1, N, 1, xg, Pf.view(N, 1, K), 1, xg, 1, K, Dp).view(K, D + 1, D + 1)
    finally:
        _lib.FORCE_SIMT = old
    assert float((G.double() - Gref).abs().max() / Gref.abs().max()) <= 1e-5
    asym = float((G - G.transpose(-1, -2)).abs().max() / G.abs().max())
    assert asym == 0.0 if not simt else asym < 1e-6        # the pair-GEMM kernel is symmetric by construction

    def scat(G):   # centred scatter: the cancellation-sensitive quantity the NIW update forms from the blocks
        G = G.double()
        return G[:, :D, :D] - G[:, :D, D:] @ G[:, D:, :D] / G[:, D:, D:].clamp_min(1e-30)
    s_ref = scat(Gref)
    keep = Gref[:, D, D] > 10.0                                            # components that own some mass
    err = (scat(G) - s_ref).flatten(1).norm(dim=1) / s_ref.flatten(1).norm(dim=1).clamp_min(1e-30)
    assert not bool(keep.any()) or float(err[keep].max()) <= 1e-4      # (N < 10 K: no component owns enough mass to ask)


def _units(D, decades, seed=7):
    g = torch.Generator().manual_seed(seed)
    return 10.0 ** ((torch.rand(D, generator=g) * 2 - 1) * decades)


@pytest.mark.parametrize("decades", [1.0, 3.0, 5.0])
def test_estep_features_in_different_units(decades):
    """Features whose units differ by up to 10^(2 decades): the split-precision operands are rescaled per feature (from the
    whitening factors), per sample row and per component, so the logits keep fp32 accuracy."""
    N, d0, K = 6000, 64, 64
    z, z0, z1, W, m, cst, Dp, L = _problem(N, d0, 0, K, seed=3, fscale=_units(d0, decades))
    xg = torch.zeros(1, dtype=torch.int32, device=DEV)
    lz = torch.logsumexp(L, -1)
    P = (L - lz[:, None]).exp()
    lg = _lib.estep(z0, z1, N, 1, xg, W, m, cst, 1, K, Dp, 0).view(N, K)
    p, lzn, NA, lZ = _lib.estep(z0, z1, N, 1, xg, W, m, cst, 1, K, Dp, 1)
    scale = float(L.abs().max())
    assert float((lg.double() - L).abs().max()) <= 2e-6 * scale
    assert float((lzn.view(N).double() - lz).abs().max()) <= 2e-6 * scale
    assert float((p.view(N, K).double() - P).abs().max()) <= 2e-3


@pytest.mark.parametrize("decades", [1.0, 3.0, 5.0])
def test_gram_features_in_different_units(decades):
    N, d0, K = 6000, 64, 64
    u = _units(d0, decades)
    z, z0, z1, W, m, cst, Dp, L = _problem(N, d0, 0, K, seed=4, fscale=u)
    xg = torch.zeros(1, dtype=torch.int32, device=DEV)
    Pf = (L - torch.logsumexp(L, -1)[:, None]).exp().float().contiguous()
    zt = torch.cat([z.double(), torch.ones(N, 1, device=DEV, dtype=torch.float64)], 1)
    Gref = torch.einsum("nk,ni,nj->kij", Pf.double(), zt, zt)
    G = _lib.gram(z0, z1, N, 1, xg, Pf.view(N, 1, K), 1, xg, 1, K, Dp).view(K, d0 + 1, d0 + 1)
    s = torch.cat([u.double(), torch.ones(1, dtype=torch.float64)]).to(DEV)
    En = (G.double() - Gref) / (s[:, None] * s[None, :])              # every entry in its own units
    Rn = Gref / (s[:, None] * s[None, :])
    assert float(En.abs().max() / Rn.abs().max()) <= 1e-5


def test_gram_weights_beyond_fp16_range_use_the_tf32_path():
    """Weights above the fp16 window (responsibilities never are) must not lose accuracy: the device-side flag routes the
    call to the TF32 kernel."""
    N, d0, K = 5000, 64, 64
    z, z0, z1, W, m, cst, Dp, L = _problem(N, d0, 0, K, seed=5)
    xg = torch.zeros(1, dtype=torch.int32, device=DEV)
    Pf = (37.5 * (L - torch.logsumexp(L, -1)[:, None]).exp()).float().contiguous()
    zt = torch.cat([z.double(), torch.ones(N, 1, device=DEV, dtype=torch.float64)], 1)
    Gref = torch.einsum("nk,ni,nj->kij", Pf.double(), zt, zt)
    G = _lib.gram(z0, z1, N, 1, xg, Pf.view(N, 1, K), 1, xg, 1, K, Dp).view(K, d0 + 1, d0 + 1)
    assert float((G.double() - Gref).abs().max() / Gref.abs().max()) <= 1e-5


def _per_component_relerr(G, Gref, min_mass=4.0):
    """max over components that own some mass of ||G_k - Gref_k||_F / ||Gref_k||_F, and the same for each component's
    diagonal entry by entry (every feature in that component's own scale)."""
    D = Gref.shape[-1] - 1
    keep = Gref[:, D, D] > min_mass
    E = (G.double() - Gref)
    fro = E.flatten(1).norm(dim=1) / Gref.flatten(1).norm(dim=1).clamp_min(1e-300)
    dg = E.diagonal(dim1=-1, dim2=-2).abs() / Gref.diagonal(dim1=-1, dim2=-2).abs().clamp_min(1e-300)
    return float(fro[keep].max()), float(dg[keep].max())


@pytest.mark.parametrize("case", ["outlier", "tight_cluster", "one_hot", "plain"])
def test_gram_data_outside_the_fp16_window(case):
    """One feature scale per COLUMN cannot cover every data set in fp16's window: a single huge outlier in a column pushes
    every ordinary sample ~20 binades under that column's scale, and a tight cluster inside a wide data range sits far below
    every column's scale.  The fp16 kernel's products for those samples fall into fp16 subnormals (or flush to zero: the
    covariance of the components that own them would collapse); the first reduce detects a component whose mean square in
    a feature is below the resolvable floor and the TF32 kernel recomputes.  Gate: PER-COMPONENT relative error, each
    diagonal entry in its own scale — a global-maximum norm cannot see the damage."""
    N, d0, K = 6000, 64, 16
    dev = torch.device(DEV)
    g = torch.Generator(device=dev).manual_seed(11)
    lab = torch.randint(K, (N,), generator=g, device=dev)
    mu = torch.randn(K, d0, generator=g, device=dev)
    z = mu[lab] + 0.5 * torch.randn(N, d0, generator=g, device=dev)
    if case == "outlier":
        z[17, 5] = 1.0e6                                       # one wild value in column 5
    elif case == "tight_cluster":
        own = lab == 3
        z[own] = 1e-5 * (1.0 + 0.1 * torch.randn(int(own.sum()), d0, generator=g, device=dev))    # values ~1e-5 of the range
    elif case == "one_hot":
        # exact zeros are exact in any format: columns without small NONZERO values are exempt from the check (a component
        # that never sees feature f has mean z_f^2 = 0, which must not be mistaken for "below the resolvable floor")
        z[:, :K] = torch.nn.functional.one_hot(lab, K).float()
    P = torch.full((N, K), 1e-4, device=dev)
    P[torch.arange(N, device=dev), lab] = 1.0 - 1e-4 * (K - 1)
    P = P.contiguous()
    xg = torch.zeros(1, dtype=torch.int32, device=DEV)
    zt = torch.cat([z.double(), torch.ones(N, 1, device=DEV, dtype=torch.float64)], 1)
    Gref = torch.einsum("nk,ni,nj->kij", P.double(), zt, zt)
    G = _lib.gram(z.contiguous(), None, N, 1, xg, P.view(N, 1, K), 1, xg, 1, K, _lib.pad_dim(d0)).view(K, d0 + 1, d0 + 1)
    fro, dg = _per_component_relerr(G, Gref)
    assert fro <= 1e-5, (case, fro)
    assert dg <= 2e-5, (case, dg)


def test_gram_sample_image_is_cached_per_data_set_and_invalidated_by_edits():
    """K3's sample image (column maxima + transposed, pre-scaled chunks) is made once per data set: a second call on the
    same rows reuses it (bit-identical Gram), an in-place edit of the rows (torch's version counter) or other rows at the
    same size do not."""
    N, d0, d1, K = 5000, 32, 32, 64
    z, z0, z1, W, m, cst, Dp, L = _problem(N, d0, d1, K, seed=12)
    xg = torch.zeros(1, dtype=torch.int32, device=DEV)
    P = (L - torch.logsumexp(L, -1)[:, None]).exp().float().contiguous().view(N, 1, K)
    key = _lib._rpack_key(P.device)
    _lib._zpack_cache.clear()
    n0 = _lib.LAUNCHES
    G1 = _lib.gram(z0, z1, N, 1, xg, P, 1, xg, 1, K, Dp).clone()
    n1 = _lib.LAUNCHES
    assert key in _lib._zpack_cache
    G2 = _lib.gram(z0, z1, N, 1, xg, P, 1, xg, 1, K, Dp).clone()
    n2 = _lib.LAUNCHES
    assert torch.equal(G1, G2)
    assert (n1 - n0) - (n2 - n1) == 3                                    # the second call skipped the image kernels
                                                                         # (column maxima of z0 and of z1, transpose)
    old = _lib.ZCACHE
    _lib.ZCACHE = 0
    try:
        G0 = _lib.gram(z0, z1, N, 1, xg, P, 1, xg, 1, K, Dp).clone()     # image made inside the call's workspace
    finally:
        _lib.ZCACHE = old
    assert torch.equal(G0, G1)
    z0.mul_(3.0)                                                         # in-place edit: the cached image is stale
    G3 = _lib.gram(z0, z1, N, 1, xg, P, 1, xg, 1, K, Dp)
    zt = torch.cat([z0.double(), z1.double(), torch.ones(N, 1, device=DEV, dtype=torch.float64)], 1)
    Gref = torch.einsum("nk,ni,nj->kij", P.view(N, K).double(), zt, zt)
    assert float((G3.double() - Gref).abs().max() / Gref.abs().max()) <= 1e-5
    z0b = (z0 * 0.5).contiguous()                                        # other rows, same shape
    G4 = _lib.gram(z0b, z1, N, 1, xg, P, 1, xg, 1, K, Dp)
    zt = torch.cat([z0b.double(), z1.double(), torch.ones(N, 1, device=DEV, dtype=torch.float64)], 1)
    Gref = torch.einsum("nk,ni,nj->kij", P.view(N, K).double(), zt, zt)
    assert float((G4.double() - Gref).abs().max() / Gref.abs().max()) <= 1e-5


@pytest.mark.parametrize("N,d0,d1,K", [(5000, 64, 0, 256), (3001, 32, 32, 64), (2049, 16, 0, 132), (4099, 16, 16, 32)])
def test_estep_hands_presplit_weights_to_gram(N, d0, d1, K):
    """K2 (mode 1) leaves the responsibilities pre-split for K3; the Gram from those images must equal (bit for bit) the
    Gram K3 computes when it splits the same p itself, and an edited p must not use stale images."""
    z, z0, z1, W, m, cst, Dp, L = _problem(N, d0, d1, K, seed=6)
    xg = torch.zeros(1, dtype=torch.int32, device=DEV)
    p, lzn, NA, lZ = _lib.estep(z0, z1, N, 1, xg, W, m, cst, 1, K, Dp, 1)
    key = _lib._rpack_key(p.device)
    assert key in _lib._rpack_rec                                       # the hand-over happened
    G1 = _lib.gram(z0, z1, N, 1, xg, p, 1, xg, 1, K, Dp).clone()
    G2 = _lib.gram(z0, z1, N, 1, xg, p.clone(), 1, xg, 1, K, Dp).clone()       # a copy: K3 splits it itself
    assert torch.equal(G1, G2)
    p.mul_(0.5)                                                         # in-place edit: version counter moves on
    G3 = _lib.gram(z0, z1, N, 1, xg, p, 1, xg, 1, K, Dp)
    assert float((G3.double() - 0.5 * G2.double()).abs().max() / G2.abs().max()) <= 1e-5


def test_tf32_operand_variants_still_agree():
    """The TF32-split kernels stay in the library (VBMP_ESTEP_PREC / VBMP_GRAM_PREC = tf32, read once per process, and the
    Gram's on-device fallback): run them in a fresh interpreter against the same fp64 restatement."""
    import os, subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, torch
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import test_cuda_kernels as T
        from pyvbmp_b200 import _lib
        N, d0, K = 5000, 64, 64
        z, z0, z1, W, m, cst, Dp, L = T._problem(N, d0, 0, K, seed=8)
        xg = torch.zeros(1, dtype=torch.int32, device=T.DEV)
        lg = _lib.estep(z0, z1, N, 1, xg, W, m, cst, 1, K, Dp, 0).view(N, K)
        p, lzn, NA, lZ = _lib.estep(z0, z1, N, 1, xg, W, m, cst, 1, K, Dp, 1)
        scale = float(L.abs().max())
        assert float((lg.double() - L).abs().max()) <= 2e-6 * scale
        P = (L - torch.logsumexp(L, -1)[:, None]).exp()
        assert float((p.view(N, K).double() - P).abs().max()) <= 2e-3
        zt = torch.cat([z.double(), torch.ones(N, 1, device=T.DEV, dtype=torch.float64)], 1)
        Gref = torch.einsum("nk,ni,nj->kij", p.view(N, K).double(), zt, zt)
        G = _lib.gram(z0, z1, N, 1, xg, p, 1, xg, 1, K, Dp).view(K, d0 + 1, d0 + 1)
        assert float((G.double() - Gref).abs().max() / Gref.abs().max()) <= 1e-5
        print("tf32 variants ok")
    """) % (os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, VBMP_ESTEP_PREC="tf32", VBMP_GRAM_PREC="tf32")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "tf32 variants ok" in out.stdout, out.stderr[-2000:]


def test_gram_is_deterministic_run_to_run():
    z, z0, z1, W, m, cst, Dp, L = _problem(40000, 64, 0, 64, seed=2)
    xg = torch.zeros(1, dtype=torch.int32, device=DEV)
    P = (L - torch.logsumexp(L, -1)[:, None]).exp().float().contiguous()
    a = _lib.gram(z0, None, 40000, 1, xg, P.view(40000, 1, 64), 1, xg, 1, 64, Dp).clone()
    b = _lib.gram(z0, None, 40000, 1, xg, P.view(40000, 1, 64), 1, xg, 1, 64, Dp).clone()
    assert torch.equal(a, b)
    r1 = _lib.estep(z0, None, 40000, 1, xg, W, m, cst, 1, 64, Dp, 1)
    r1 = [t.clone() for t in r1]
    r2 = _lib.estep(z0, None, 40000, 1, xg, W, m, cst, 1, 64, Dp, 1)
    for u, v in zip(r1, r2):
        assert torch.equal(u, v)


@pytest.mark.parametrize("T,S,G,K,ptemp", [(50, 37, 1, 5, 1.0), (128, 64, 1, 32, 1.0), (17, 12, 3, 24, 2.0), (1, 9, 1, 4, 1.0)])
def test_hmm_forward_backward_kernel(T, S, G, K, ptemp):
    """K6 against the torch restatement of models/HMM.py:72-105 in fp64 (same recursion, batched ops)."""
    import pyvbmp_b200 as V
    g = torch.Generator(device=DEV).manual_seed(T + K)
    rest = (S, G) if G > 1 else (S,)
    logits = 3.0 * torch.randn((T,) + rest + (K,), generator=g, device=DEV) - 5.0
    tr = torch.log_softmax(torch.randn((G, K, K) if G > 1 else (K, K), generator=g, device=DEV), -1)
    init = torch.log_softmax(torch.randn((G, K) if G > 1 else (K,), generator=g, device=DEV), -1)

    class _D:        # stand-ins for the Dirichlet nodes: only loggeomean() is used
        def __init__(self, v): self.v = v
        def loggeomean(self): return self.v
    h = V.HMM.__new__(V.HMM)
    h.batch_shape = (G,) if G > 1 else ()
    h.event_shape = (K,)
    h.ptemp = ptemp
    h.transition, h.initial = _D(tr), _D(init)
    p, SEzz, SEz0, logZ = h.forward_backward_logits(logits.clone())
    h.transition, h.initial = _D(tr.double()), _D(init.double())
    pr, SEzzr, SEz0r, logZr = h._forward_backward_torch(logits.double(), tr.double(), init.double())
    assert p.shape == pr.shape and SEzz.shape == SEzzr.shape and SEz0.shape == SEz0r.shape and logZ.shape == logZr.shape
    assert float((p.double() - pr).abs().max()) < 2e-5
    assert float((SEzz.double() - SEzzr).abs().max()) < 1e-4 * max(1.0, float(SEzzr.abs().max()))
    assert float((SEz0.double() - SEz0r).abs().max()) < 2e-5
    assert float(((logZ.double() - logZr).abs() / logZr.abs().clamp_min(1.0)).max()) < 1e-5
    assert bool((p.argmax(-1) == pr.argmax(-1)).float().mean() > 0.999)


def test_misaligned_views_are_rebased():
    """A contiguous view that starts at a 4-byte offset must not reach the 16-byte vector / TMA accesses as is."""
    import pyvbmp_b200 as V
    N, K, d = 5000, 8, 64
    g = torch.Generator(device=DEV).manual_seed(4)
    flat = torch.randn(N * d + 1, generator=g, device=DEV)
    Xa = flat[1:].view(N, d)                      # data_ptr % 16 == 4
    Xb = Xa.clone()
    assert Xa.data_ptr() % 16 != 0 and Xb.data_ptr() % 16 == 0
    outs = []
    for X in (Xa, Xb):
        torch.manual_seed(1)
        m = V.GaussianMixtureModel(K, d)
        m.initialize(Xb.cpu()[:512])
        m.to(DEV)
        m.update(X, 2)
        outs.append(m)
    assert torch.equal(outs[0].p, outs[1].p)
    assert float(outs[0].ELBO_last) == float(outs[1].ELBO_last)


@pytest.mark.parametrize("N,K,n,with_base", [(5000, 64, 32, True), (3001, 8, 16, True), (777, 5, 12, False), (1025, 20, 7, True),
                                              (300, 3, 1, False), (4000, 40, 32, False), (2000, 100, 16, True), (3, 33, 32, True)])
def test_moe_moments_kernel(N, K, n, with_base):
    """vbmp_moe_moments (the per-sample part of MixtureofLinearTransforms.predict, transforms/MixtureofLinearTransforms.py:
    100-106) against an fp64 evaluation: the tensor-core SYRK kernel (n = 16, 32; K that needs zero-filled rows in the last
    staged chunk), the 4 x 8 register-tiled kernel (other n % 4 == 0) and the row-per-lane one."""
    g = torch.Generator(device=DEV).manual_seed(N + n)
    mean = (torch.randn(N, K, n, generator=g, device=DEV) * 1.3 + 0.4).contiguous()
    p = torch.softmax(1.5 * torch.randn(N, K, generator=g, device=DEV), -1).contiguous()
    base = None
    if with_base:
        A = torch.randn(N, n, n, generator=g, device=DEV) * 0.2
        base = (A @ A.transpose(-1, -2)).contiguous()
    mu, Sig = _lib.moe_moments(mean, p, base, N, K, n)
    md, pd = mean.double(), p.double()
    mu_ref = torch.einsum("nk,nki->ni", pd, md)
    S_ref = torch.einsum("nk,nki,nkj->nij", pd, md, md) - mu_ref.unsqueeze(-1) * mu_ref.unsqueeze(-2)
    if with_base:
        S_ref = S_ref + base.double()
    assert float((mu.double() - mu_ref).abs().max()) <= 1e-5 * float(mu_ref.abs().max())
    assert float((Sig.double() - S_ref).abs().max()) <= 2e-5 * float(S_ref.abs().max())
    if with_base:      # in place: Sigma may alias base
        b2 = base.clone()
        mu2, Sig2 = _lib.moe_moments(mean, p, b2, N, K, n, Sigma=b2)
        assert torch.equal(Sig2, Sig) and torch.equal(mu2, mu)


def test_torch_custom_ops_match_the_binding():
    """torch.ops.vbmp.* are the same kernels as the ctypes binding the mirrors call."""
    import pyvbmp_b200.ops  # noqa: F401
    N, d0, d1, K = 3001, 32, 32, 64
    z, z0, z1, W, m, cst, Dp, L = _problem(N, d0, d1, K, seed=21)
    xg = torch.zeros(1, dtype=torch.int32, device=DEV)
    lg = torch.ops.vbmp.estep_logits(z0, z1, W, m, cst)
    assert torch.equal(lg, _lib.estep(z0, z1, N, 1, xg, W, m, cst, 1, K, Dp, 0).view(N, K))
    p, lzn, NA, lZ = torch.ops.vbmp.estep_assign(z0, z1, W, m, cst)
    p2, lzn2, NA2, lZ2 = _lib.estep(z0, z1, N, 1, xg, W, m, cst, 1, K, Dp, 1)
    assert torch.equal(p, p2.view(N, K)) and torch.equal(NA, NA2.view(K)) and torch.equal(lZ, lZ2.view(()))
    G = torch.ops.vbmp.gram(z0, z1, p, False)
    assert torch.equal(G, _lib.gram(z0, z1, N, 1, xg, p2.clone(), 1, xg, 1, K, Dp).view(K, d0 + d1 + 1, d0 + d1 + 1))
    # the widened rows: products with the flattened covariances, mixture-of-experts moments
    g = torch.Generator(device=DEV).manual_seed(5)
    S = torch.randn(N, 256, generator=g, device=DEV)
    L = torch.randn(256, K, generator=g, device=DEV)
    assert torch.equal(torch.ops.vbmp.wsum(p, S), _lib.wsum(p, S))
    assert torch.equal(torch.ops.vbmp.rowterm(S, L, lg, -0.5), _lib.rowterm(S, L, C=lg.clone(), alpha=-0.5, accumulate=True))
    Bm, bias = torch.randn(d0, 512, generator=g, device=DEV), torch.randn(512, generator=g, device=DEV)
    assert torch.equal(torch.ops.vbmp.rowgemm(z0, Bm, bias), _lib.rowgemm(z0, Bm, bias=bias))
    mean = torch.randn(N, K, 16, generator=g, device=DEV)
    mu, Sig = torch.ops.vbmp.moe_moments(mean, p, None)
    mu2, Sig2 = _lib.moe_moments(mean, p, None, N, K, 16)
    assert torch.equal(mu, mu2) and torch.equal(Sig, Sig2)


@pytest.mark.parametrize("N,Kd,M,bias,acc", [(5000, 32, 2048, True, False), (3001, 64, 1024, False, False),
                                              (4100, 1024, 64, False, True), (777, 7, 13, True, True), (129, 33, 100, True, False)])
def test_rowgemm_kernel(N, Kd, M, bias, acc):
    """vbmp_rowgemm (3 x TF32 products on the warp-level tensor-core path) against an fp64 product: fp32-grade accuracy on
    aligned and ragged shapes, with bias and with accumulation into the output."""
    g = torch.Generator(device=DEV).manual_seed(N + Kd + M)
    A = torch.randn(N, Kd, generator=g, device=DEV) * 1.7 + 0.3
    B = torch.randn(Kd, M, generator=g, device=DEV)
    b = torch.randn(M, generator=g, device=DEV) if bias else None
    C0 = torch.randn(N, M, generator=g, device=DEV) if acc else None
    ref = A.double() @ B.double()
    if bias:
        ref = ref + b.double()
    if acc:
        ref = ref + C0.double()
    out = _lib.rowgemm(A, B, bias=b, out=None if not acc else C0.clone(), accumulate=acc)
    scale = float((A.abs().double() @ B.abs().double()).max())
    assert float((out.double() - ref).abs().max()) <= 2e-6 * scale
    # a strided view as A (rows of a wider matrix) and as the output
    if Kd >= 8:
        wide = torch.randn(N, Kd + 12, generator=g, device=DEV)
        out2 = torch.zeros(N, M + 4, device=DEV)
        _lib.rowgemm(wide[:, 4:4 + Kd], B, out=out2[:, :M])
        ref2 = wide[:, 4:4 + Kd].double() @ B.double()
        assert float((out2[:, :M].double() - ref2).abs().max()) <= 2e-6 * float((wide.abs().double()[:, 4:4 + Kd] @ B.abs().double()).max())
        assert float(out2[:, M:].abs().max()) == 0.0


@pytest.mark.parametrize("N,K,F,lds", [(5000, 64, 1024, 1024), (4100, 256, 100, 128), (2048, 8, 16, 16), (70000, 32, 289, 292),
                                       (3000, 128, 1089, 1092)])
def test_wsum_kernel(N, K, F, lds):
    """vbmp_wsum (weighted column sums over the sample axis on the Gram kernel's "lin" mode, TF32 split) against an fp64
    product: fp32-grade accuracy per entry, column blocks that do not fill the last TMA box, a row stride wider than F,
    and bit-reproducibility."""
    g = torch.Generator(device=DEV).manual_seed(N + K + F)
    p = torch.softmax(3.0 * torch.randn(N, K, generator=g, device=DEV), dim=-1)
    wide = torch.randn(N, lds, generator=g, device=DEV) * 1.3 + 0.5
    S = wide[:, :F]
    assert _lib.wsum_supported(N, K, F, S.stride(0))
    out = _lib.wsum(p, S)
    ref = p.double().t() @ S.double()
    scale = p.double().t() @ S.abs().double()                  # per entry: what the products could add up to
    assert float(((out.double() - ref).abs() / scale).max()) <= 4e-6
    assert torch.equal(out, _lib.wsum(p, S))


@pytest.mark.parametrize("N,F,K,lda,acc", [(5000, 1024, 64, 1024, True), (4100, 256, 8, 256, False), (129, 36, 5, 40, True),
                                           (70000, 1024, 32, 1024, False), (3000, 100, 200, 104, True), (20000, 32, 64, 32, False)])
def test_rowterm_kernel(N, F, K, lda, acc):
    """vbmp_rowterm (tcgen05: A from HBM through registers into tensor memory, 3-term TF32 split) against an fp64 product:
    fp32-grade accuracy, a ragged last row tile, feature counts that do not fill the last chunk, K that is not a multiple of
    16, a row-strided A, alpha and accumulation."""
    g = torch.Generator(device=DEV).manual_seed(N + F + K)
    wide = torch.randn(N, lda, generator=g, device=DEV) * 1.3 + 0.4
    A = wide[:, :F]
    B = torch.randn(F, K, generator=g, device=DEV)
    C0 = torch.randn(N, K, generator=g, device=DEV) if acc else None
    assert _lib.rowterm_supported(N, F, K, A.stride(0))
    out = _lib.rowterm(A, B, C=None if not acc else C0.clone(), alpha=-0.5, accumulate=acc)
    ref = -0.5 * (A.double() @ B.double())
    if acc:
        ref = ref + C0.double()
    scale = 0.5 * (A.abs().double() @ B.abs().double()) + (C0.abs().double() if acc else 0.0)
    assert float(((out.double() - ref).abs() / scale).max()) <= 2e-6


@pytest.mark.parametrize("N,Kd,M,bias,acc", [(128, 8, 64, False, False), (40000, 33, 2048, False, True), (20000, 63, 1000, True, False),
                                              (19000, 17, 204, True, True), (70000, 64, 1024, False, False)])
def test_rowgemm_short_reduction_wide_output(N, Kd, M, bias, acc):
    """The tcgen05 kernel behind vbmp_rowgemm_ex (A tile in tensor memory, packed 128-column B chunks, output transposed
    through shared memory): reduction lengths that need zero padding to 8, a bias row that is the 64th reduction index,
    a last column chunk that is not full, a ragged last row tile, several tiles per CTA, accumulation."""
    g = torch.Generator(device=DEV).manual_seed(N + Kd + M)
    A = torch.randn(N, Kd, generator=g, device=DEV) * 1.7 + 0.3
    B = torch.randn(Kd, M, generator=g, device=DEV)
    b = torch.randn(M, generator=g, device=DEV) if bias else None
    C0 = torch.randn(N, M, generator=g, device=DEV) if acc else None
    ref = A.double() @ B.double()
    if bias:
        ref = ref + b.double()
    if acc:
        ref = ref + C0.double()
    out = _lib.rowgemm(A, B, bias=b, out=None if not acc else C0.clone(), accumulate=acc)
    scale = float((A.abs().double() @ B.abs().double()).max())
    assert float((out.double() - ref).abs().max()) <= 2e-6 * scale


@pytest.mark.parametrize("N,K,bias,inplace", [(5000, 64, True, True), (129, 5, False, False), (3001, 700, True, True),
                                               (260, 2049, False, True), (70000, 32, True, False)])
def test_softmax_rows_kernel(N, K, bias, inplace):
    """vbmp_softmax_rows against an fp64 evaluation: responsibilities, per-row log normaliser, column sums, their
    bit-reproducibility, in place over the logits and into a separate buffer, a row-strided logits view."""
    g = torch.Generator(device=DEV).manual_seed(N + K)
    wide = 6.0 * torch.randn(N, K + 4, generator=g, device=DEV)
    lg = wide[:, :K]
    b = torch.randn(K, generator=g, device=DEV) if bias else None
    L = lg.double() + (b.double() if bias else 0.0)
    lz = torch.logsumexp(L, -1)
    P = (L - lz[:, None]).exp()
    src = lg.clone() if inplace else lg
    p, lzn, NA, lZ = _lib.softmax_rows(src, colbias=b, out=src if inplace else None)
    assert float((p.double() - P).abs().max()) <= 2e-6
    assert float((lzn.double() - lz).abs().max()) <= 2e-6 * float(L.abs().max())
    assert float(((NA.double() - P.sum(0)).abs() / P.sum(0).clamp_min(1.0)).max()) <= 1e-5
    assert abs(float(lZ.double() - lz.sum())) <= 1e-6 * abs(float(lz.sum()))
    p2, lzn2, NA2, lZ2 = _lib.softmax_rows(lg.clone(), colbias=b)
    assert torch.equal(NA, NA2) and torch.equal(lZ, lZ2) and torch.equal(p, p2)
