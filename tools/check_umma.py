"""GPU check of the tcgen05 kernels against the CUDA-core kernels and an fp64 torch reference (dev tool;
the same comparisons live in tests/test_cuda_kernels.py)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyvbmp_b200 import _lib

dev = torch.device("cuda:0")


def make(N, d0, d1, K, seed=0, spread=1.0):
    g = torch.Generator(device=dev).manual_seed(seed)
    D = d0 + d1
    Dp = _lib.pad_dim(D)
    mu = spread * torch.randn(K, D, generator=g, device=dev)
    z = mu[torch.randint(K, (N,), generator=g, device=dev)] + torch.randn(N, D, generator=g, device=dev)
    A = torch.randn(K, D, D, generator=g, device=dev) / D ** 0.5
    invU = A @ A.transpose(-1, -2) + 0.5 * torch.eye(D, device=dev)
    nu = D + 2 + 10 * torch.rand(K, generator=g, device=dev)
    lam = 1 + torch.rand(K, generator=g, device=dev)
    lp = torch.log_softmax(torch.randn(K, generator=g, device=dev), 0)
    W, m, cst, info = _lib.niw_prep(invU.contiguous(), mu.contiguous(), nu, lam, lp, K, D, Dp)
    z0 = z[:, :d0].contiguous()
    z1 = z[:, d0:].contiguous() if d1 else None
    return z, z0, z1, W, m, cst, Dp


def ref_logits(z, W, m, cst, D):
    y = torch.einsum("ni,kij->nkj", z.double(), W[:, :D, :].double()) - m.double()[None]
    return cst.double()[None] - 0.5 * (y * y).sum(-1)


def check(N, d0, d1, K, seed=0):
    D = d0 + d1
    z, z0, z1, W, m, cst, Dp = make(N, d0, d1, K, seed)
    xg = torch.zeros(1, dtype=torch.int32, device=dev)
    L = ref_logits(z, W, m, cst, D)
    lz = torch.logsumexp(L, -1)
    P = (L - lz[:, None]).exp()
    res = {}
    for simt in (1, 0):
        _lib.FORCE_SIMT = simt
        lg = _lib.estep(z0, z1, N, 1, xg, W, m, cst, 1, K, Dp, 0).view(N, K)
        p, lzn, NA, lZ = _lib.estep(z0, z1, N, 1, xg, W, m, cst, 1, K, Dp, 1)
        torch.cuda.synchronize()
        p = p.view(N, K)
        e_l = float((lg.double() - L).abs().max())
        e_p = float((p.double() - P).abs().max())
        e_z = float((lzn.view(N).double() - lz).abs().max())
        e_na = float(((NA.view(K).double() - P.sum(0)).abs() / P.sum(0).abs().clamp_min(1)).max())
        e_lZ = float((lZ.double().sum() - lz.sum()).abs() / lz.sum().abs())
        am = int((p.argmax(-1) != P.argmax(-1)).sum())
        res[simt] = (e_l, e_p, e_z, e_na, e_lZ, am)
        print(f"estep N={N} d=({d0},{d1}) K={K} {'simt' if simt else 'umma'}: |dlogit|={e_l:.2e} |dp|={e_p:.2e} "
              f"|dlogZn|={e_z:.2e} NA rel={e_na:.2e} logZ rel={e_lZ:.2e} argmax mismatches={am} (|logit| max {float(L.abs().max()):.1f})")
    # gram
    Pf = P.float().contiguous()
    zt = torch.cat([z.double(), torch.ones(N, 1, device=dev, dtype=torch.float64)], 1)
    Gref = torch.einsum("nk,ni,nj->kij", P, zt, zt) if N * K * (D + 1) ** 2 < 3e9 else None
    if Gref is None:
        Gref = torch.zeros(K, D + 1, D + 1, device=dev, dtype=torch.float64)
        for a in range(0, N, 8192):
            Gref += torch.einsum("nk,ni,nj->kij", P[a:a + 8192], zt[a:a + 8192], zt[a:a + 8192])
    for simt in (1, 0):
        _lib.FORCE_SIMT = simt
        G = _lib.gram(z0, z1, N, 1, xg, Pf, 1, xg, 1, K, Dp).view(K, D + 1, D + 1)
        torch.cuda.synchronize()
        rel = float((G.double() - Gref).abs().max() / Gref.abs().max())
        # centred scatter (the cancellation-sensitive quantity): S_k = Gxx - Gx Gx^T / N_k
        def scat(G):
            G = G.double()
            return G[:, :D, :D] - G[:, :D, D:] @ G[:, D:, :D] / G[:, D:, D:].clamp_min(1e-30)
        s_ref = scat(Gref)
        srel = float(((scat(G) - s_ref).flatten(1).norm(dim=1) / s_ref.flatten(1).norm(dim=1).clamp_min(1e-30)).max())
        print(f"gram  N={N} d=({d0},{d1}) K={K} {'simt' if simt else 'umma'}: max rel err {rel:.2e}, centred scatter rel (worst comp) {srel:.2e}")
    _lib.FORCE_SIMT = 0


if __name__ == "__main__":
    cases = [(5000, 64, 0, 256), (70000, 64, 0, 256), (3001, 32, 32, 64), (4099, 16, 16, 32), (2500, 16, 0, 8), (6000, 48, 0, 20)]
    if len(sys.argv) > 1:
        cases = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
    for c in cases:
        check(*c)
    print("check_umma done")
