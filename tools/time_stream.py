"""Sweep the streamed-chunk size of GaussianMixtureModel.update(X_host) at cfg2 (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyvbmp_b200 as V
from pyvbmp_b200.mixture import Mixture
dev = torch.device("cuda:0")
N, d, K = 1 << 22, 64, 256
g = torch.Generator().manual_seed(0)
mu = 3 * torch.randn(K, d, generator=g)
Xh = torch.empty(N, d).pin_memory()
for a in range(0, N, 1 << 18):
    Xh[a:a + (1 << 18)] = mu[torch.randint(K, (1 << 18,), generator=g)] + torch.randn(1 << 18, d, generator=g)
torch.manual_seed(0)
m = V.GaussianMixtureModel(K, d).to(dev)
m.initialize(Xh[:65536].to(dev))
for rows, first in ((1 << 19, 1 << 15), (1 << 18, 1 << 15), (1 << 17, 1 << 15), (1 << 19, 1 << 19), (1 << 18, 1 << 18)):
    Mixture.STREAM_ROWS, Mixture.STREAM_FIRST = rows, first
    m._stream_state = None
    for _ in range(2): m.update(Xh, 1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        m.update(Xh, 1)
        float(m.ELBO_last)
    b.record(); b.synchronize()
    print(f"STREAM_ROWS={rows} STREAM_FIRST={first}: {a.elapsed_time(b) / 5:.2f} ms per iteration")
