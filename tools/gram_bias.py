"""Bias / error of the tcgen05 Gram kernel vs fp64 for random and clustered weights (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyvbmp_b200 import _lib
dev = torch.device("cuda:0")
N, K, d = 1 << 18, 64, 32
g = torch.Generator(device=dev).manual_seed(5)
X = torch.randn(N, d, generator=g, device=dev) * 1.5 + 0.5
r = torch.rand(N, K, generator=g, device=dev)
xg = torch.zeros(1, dtype=torch.int32, device=dev)
G = _lib.gram(X.view(N, 1, d), None, N, 1, xg, r.view(N, 1, K), 1, xg, 1, K, 32).view(K, d + 1, d + 1).double()
Z1 = torch.cat([X, torch.ones(N, 1, device=dev)], -1).double()
ref = torch.einsum("nk,ni,nj->kij", r[:, :4].double(), Z1, Z1)
rel = (G[:4] - ref) / ref.abs().clamp_min(1e-30)
big = ref.abs() > 0.1 * ref.abs().max()
def scat(G):
    return G[:, :d, :d] - G[:, :d, d:] @ G[:, d:, :d] / G[:, d:, d:]
S, Sr = scat(G[:4]), scat(ref)
print(f"FL={os.environ.get('VBMP_GRAM_FL','dflt')} split={os.environ.get('VBMP_GRAM_SPLIT','0')}: mean rel err (large entries) {float(rel[big].mean()):+.3e}, "
      f"std {float(rel[big].std()):.3e}, max |rel| {float(rel[big].abs().max()):.3e}; centred scatter rel {float((S-Sr).norm()/Sr.norm()):.3e}")
