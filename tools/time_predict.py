"""Time MixtureofLinearTransforms.predict on held-out inputs (dev tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyvbmp_b200 as V
from pyvbmp_b200 import _lib
dev = torch.device("cuda:0")
n, p, K = 32, 32, 64
N = int(os.environ.get("TP_N", 1 << 20))
g = torch.Generator(device=dev).manual_seed(0)
torch.manual_seed(0)
m = V.MixtureofLinearTransforms(n, p, K).to(dev)
X = torch.randn(N, p, 1, generator=g, device=dev)
Wt = torch.randn(K, n, p, generator=g, device=dev) / p ** 0.5
z = torch.randint(K, (N,), generator=g, device=dev)
Y = (torch.einsum("nij,nj->ni", Wt[z], X[..., 0]) + 0.1 * torch.randn(N, n, generator=g, device=dev)).unsqueeze(-1)
m.raw_update(X, Y, iters=2)
def run():
    return m.predict(X)
for _ in range(2): run()
torch.cuda.synchronize()
# THE per-call figure: CUDA events around three calls, nothing else in the region (outputs dropped as a caller's loop would)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3): run()
b.record(); b.synchronize()
t = a.elapsed_time(b) / 3
t0 = time.perf_counter()
for _ in range(3): run()
torch.cuda.synchronize()
tw = (time.perf_counter() - t0) / 3 * 1e3
print(f"MoLT.predict N={N} p={p} n={n} K={K}: {t:.2f} ms per call on the device ({tw:.2f} ms wall clock; "
      f"{N * K / t / 1e6:.2f}e9 sample*component evaluations/s)")
# Share of the gate kernel, from per-C-ABI-call events.  This loop is NOT a timing of predict: creating and recording an event
# pair around every library call, and holding the previous call's 4.3 GB of outputs while the next call allocates, put
# 1 - 35 ms of host / allocator time into it depending on the box (the kernels inside are the same).
_lib.profile_begin(512)
for _ in range(3): pY, pr = run()
torch.cuda.synchronize()
ke = sum(x.elapsed_time(y) for x, y in _lib.PROFILE.get("vbmp_estep", [])) / 3
print(f"gate probabilities (K2 kernel) {ke:.2f} ms of it; moment sums (means + base row GEMMs + moe_moments, "
      f"{N * n * n * 4 / 1e9:.2f} GB of covariances out) {t - ke:.2f} ms")

# ---- breakdown of one 64 Ki-row block: component means GEMM, base GEMM, per-sample moments kernel
W = m.W
p_in = W.p - int(W.pad_X)
M = W.mu
Mw = M[..., :p_in].reshape(K * n, p_in).t().contiguous()
Mb = M[..., -1].reshape(1, K * n)
ESf = W.EinvSigma().inverse().expand(K, n, n).reshape(K, n * n).contiguous()
rows = 1 << 16
X2 = X.view(N, p)[:rows].contiguous()
pe = pr[:rows].contiguous()
S = torch.empty(rows, n, n, device=dev)
mu = torch.empty(rows, n, device=dev)
def ev():
    return torch.cuda.Event(enable_timing=True)
ts = {}
for rep in range(3):
    e = [ev() for _ in range(4)]
    e[0].record()
    mean = _lib.rowgemm(X2, Mw, bias=Mb.reshape(-1).contiguous())
    e[1].record()
    _lib.rowgemm(pe, ESf, out=S.view(rows, n * n))
    e[2].record()
    _lib.moe_moments(mean, pe, S, rows, K, n, mu=mu, Sigma=S)
    e[3].record()
    torch.cuda.synchronize()
    ts = {"means GEMM": e[0].elapsed_time(e[1]), "base GEMM": e[1].elapsed_time(e[2]), "moe_moments": e[2].elapsed_time(e[3])}
print("per 64 Ki rows:", {k: round(v, 3) for k, v in ts.items()}, "-> x16 per 1 Mi rows:", {k: round(16 * v, 2) for k, v in ts.items()})
