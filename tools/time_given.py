"""Time MixtureofLinearTransforms.update(pX, pY) — VB-EM on Gaussian beliefs (dev tool)."""
import os, sys
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
Sx = (0.01 * torch.eye(p, device=dev)).expand(N, p, p).contiguous()
Sy = (0.01 * torch.eye(n, device=dev)).expand(N, n, n).contiguous()
pX, pY = V.MultivariateNormal_vector_format(mu=X, Sigma=Sx), V.MultivariateNormal_vector_format(mu=Y, Sigma=Sy)
for _ in range(2): m.update(pX, pY, iters=1)
torch.cuda.synchronize()
_lib.profile_begin(512)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
R = 3
for _ in range(R): m.update(pX, pY, iters=1)
b.record(); b.synchronize()
t = a.elapsed_time(b) / R
k = {name: sum(x.elapsed_time(y) for x, y in ev) / R for name, ev in _lib.PROFILE.items()}
print(f"MoLT.update(pX, pY) N={N} p={p} n={n} K={K}: {t:.2f} ms per iteration ({N * K / t / 1e6:.2f}e9 sample*component updates/s); "
      f"kernels {{{', '.join(f'{a}: {v:.2f}' for a, v in k.items())}}}; covariance products + softmax (torch) {t - sum(k.values()):.2f} ms; ELBO {float(m.ELBO_last):.6e}")
