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
from pyvbmp_b200 import _lib
for rows, first, growth in ((1 << 19, 1 << 15, 1.3), (1 << 19, 1 << 15, 1.3), (1 << 19, 1 << 15, 2.0), (1 << 20, 1 << 15, 1.3), (1 << 20, 1 << 16, 1.25), (1 << 19, 1 << 16, 1.3), (1 << 19, 1 << 14, 1.3)):
    Mixture.STREAM_ROWS, Mixture.STREAM_FIRST, Mixture.STREAM_GROWTH = rows, first, growth
    m._stream_state = None
    for _ in range(2): m.update(Xh, 1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        m.update(Xh, 1)
        float(m.ELBO_last)
    b.record(); b.synchronize()
    t = a.elapsed_time(b) / 5
    _lib.profile_begin(256)
    m.update(Xh, 1); float(m.ELBO_last)
    torch.cuda.synchronize()
    prof = _lib.profile_end()
    k = {name: (len(ev), round(sum(x.elapsed_time(y) for x, y in ev), 2)) for name, ev in prof.items()}
    print(f"STREAM_ROWS={rows} STREAM_FIRST={first} GROWTH={growth}: {t:.2f} ms per iteration; per-call events (count, ms): {k}; sum {sum(v[1] for v in k.values()):.2f}")
