"""Per-iteration timings of the other BASELINE.json configs on one B200 (dev tool):
cfg3 MixtureofLinearTransforms N=8M, p=n=32, K=64 (pad_X) and cfg4 ARHMM 4096 sequences x T=1024, d=16, K=32."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyvbmp_b200 as V
from pyvbmp_b200 import _lib

dev = torch.device("cuda:0")


def timed(fn, n=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.profile_begin(512)
    a.record()
    for _ in range(n):
        fn()
    b.record(); b.synchronize()
    prof = _lib.profile_end()
    ker = {k: round(sum(x.elapsed_time(y) for x, y in v) / n, 3) for k, v in prof.items()}
    return a.elapsed_time(b) / n, ker


def cfg3(N=8_388_608, p=32, n=32, K=64):
    g = torch.Generator(device=dev).manual_seed(1)
    X = torch.randn(N, p, generator=g, device=dev)
    W = torch.randn(K, n, p, generator=g, device=dev) / p ** 0.5
    b = torch.randn(K, n, generator=g, device=dev)
    z = torch.randint(K, (N,), generator=g, device=dev)
    Y = torch.empty(N, n, device=dev)
    for a in range(0, N, 1 << 20):
        e = min(a + (1 << 20), N)
        Y[a:e] = torch.einsum("nij,nj->ni", W[z[a:e]], X[a:e]) + b[z[a:e]] + 0.1 * torch.randn(e - a, n, generator=g, device=dev)
    torch.manual_seed(0)
    m = V.MixtureofLinearTransforms(n, p, K, pad_X=True).to(dev)
    Xc, Yc = X.unsqueeze(-1), Y.unsqueeze(-1)
    t, ker = timed(lambda: m.raw_update(Xc, Yc, iters=1, lr=1))
    print(f"cfg3 MoLT N={N} p={p} n={n} K={K}: {t:.2f} ms/iter -> {N * K / t / 1e6:.2f}e9 updates/s; kernels {ker}; ELBO {float(m.ELBO_last):.6e}")


def cfg4(S=4096, T=1024, d=16, K=32):
    g = torch.Generator(device=dev).manual_seed(2)
    A = 0.95 * torch.linalg.qr(torch.randn(K, d, d, generator=g, device=dev))[0]
    P = 4 * torch.eye(K, device=dev) + torch.rand(K, K, generator=g, device=dev)
    P = P / P.sum(-1, keepdim=True)
    y = torch.zeros(T + 1, S, d, device=dev)
    zt = torch.randint(K, (S,), generator=g, device=dev)
    y[0] = torch.randn(S, d, generator=g, device=dev)
    for t in range(T):
        y[t + 1] = torch.einsum("sij,sj->si", A[zt], y[t]) + 0.3 * torch.randn(S, d, generator=g, device=dev)
        zt = torch.multinomial(P[zt], 1, generator=g).squeeze(-1)
    X = y[:-1].reshape(T, S, 1, d, 1).contiguous()
    Y = y[1:].reshape(T, S, 1, d, 1).contiguous()
    torch.manual_seed(0)
    m = V.ARHMM(K, d, d).to(dev)
    t, ker = timed(lambda: m.update((X, Y), iters=1, lr=1))
    print(f"cfg4 ARHMM S={S} T={T} d={d} K={K}: {t:.2f} ms/iter -> {S * T * K / t / 1e6:.2f}e9 updates/s; kernels {ker}; ELBO {float(m.ELBO_last):.6e}")


if __name__ == "__main__":
    which = sys.argv[1:] or ["3", "4"]
    if "3" in which:
        cfg3()
    if "4" in which:
        cfg4()
