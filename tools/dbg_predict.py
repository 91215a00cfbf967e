import os, sys, time
sys.path.insert(0, '/root/repo')
import torch
import pyvbmp_b200 as V
from pyvbmp_b200 import _lib, _shapes
dev = torch.device("cuda:0")
n, p, K = 32, 32, 64
N = 1 << 20
g = torch.Generator(device=dev).manual_seed(0)
torch.manual_seed(0)
m = V.MixtureofLinearTransforms(n, p, K).to(dev)
X = torch.randn(N, p, 1, generator=g, device=dev)
Wt = torch.randn(K, n, p, generator=g, device=dev) / p ** 0.5
z = torch.randint(K, (N,), generator=g, device=dev)
Y = (torch.einsum("nij,nj->ni", Wt[z], X[..., 0]) + 0.1 * torch.randn(N, n, generator=g, device=dev)).unsqueeze(-1)
m.raw_update(X, Y, iters=2)
for _ in range(2): m.predict(X)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    m.predict(X)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14))
