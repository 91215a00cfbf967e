"""Workloads for the ncu captures (dev tool): python tools/prof_driver.py {cfg2|cfg3|cfg4|iso|d128} — a few EM iterations of
one configuration, sized like bench.py's."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyvbmp_b200 as V
from bench import synth_rows, D, K

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = torch.Generator(device=dev).manual_seed(1)
if which in ("cfg2", "iso"):
    N = 4_194_304
    X = synth_rows(N, dev, 1234)
    torch.manual_seed(0)
    m = V.GaussianMixtureModel(K, D, isotropic=(which == "iso")).to(dev)
    m.dist.mu = X[torch.randint(1 << 20, (K,)).to(dev)].clone()
    for _ in range(iters):
        m.update(X, 1)
elif which == "d128":
    N, Kc, d = 1 << 20, 256, 128
    mu = torch.randn(Kc, d, generator=g, device=dev)
    X = mu[torch.randint(Kc, (N,), generator=g, device=dev)] + torch.randn(N, d, generator=g, device=dev)
    torch.manual_seed(0)
    m = V.GaussianMixtureModel(Kc, d).to(dev)
    m.dist.mu = X[:Kc].clone()
    for _ in range(iters):
        m.update(X, 1)
elif which == "cfg3":
    N, p, n, Kc = 8_388_608, 32, 32, 64
    X = torch.randn(N, p, generator=g, device=dev)
    W = torch.randn(Kc, n, p, generator=g, device=dev) / p ** 0.5
    z = torch.randint(Kc, (N,), generator=g, device=dev)
    Y = torch.empty(N, n, device=dev)
    for a in range(0, N, 1 << 20):
        e = min(a + (1 << 20), N)
        Y[a:e] = torch.einsum("nij,nj->ni", W[z[a:e]], X[a:e]) + 0.1 * torch.randn(e - a, n, generator=g, device=dev)
    torch.manual_seed(0)
    m = V.MixtureofLinearTransforms(n, p, Kc, pad_X=True).to(dev)
    for _ in range(iters):
        m.raw_update(X.unsqueeze(-1), Y.unsqueeze(-1), iters=1)
elif which == "cfg4":
    S, T, d, Kc = 4096, 1024, 16, 32
    y = torch.randn(T + 1, S, d, generator=g, device=dev).cumsum(0) * 0.1
    X = y[:-1].reshape(T, S, 1, d, 1).contiguous()
    Y = y[1:].reshape(T, S, 1, d, 1).contiguous()
    torch.manual_seed(0)
    m = V.ARHMM(Kc, d, d).to(dev)
    for _ in range(iters):
        m.update((X, Y), iters=1)
torch.cuda.synchronize()
print("done", which)
