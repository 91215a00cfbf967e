"""Where a launch-bound EM iteration goes (BASELINE.json configs[0]: two-moons N=10 000, d=2, K=20): host profile of
GaussianMixtureModel.update, kernel launches per iteration, GPU time of the launches (dev tool)."""
import cProfile, math, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyvbmp_b200 as V
from pyvbmp_b200 import _lib
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(5)
x = torch.linspace(-math.pi / 2, math.pi / 2, 5000)
X = torch.cat([torch.stack([torch.sin(x), torch.cos(x) - 0.25], -1), torch.stack([torch.sin(x) + 1.0, -torch.cos(x) + 0.25], -1)], 0)
X = X + 0.05 * torch.randn(X.shape, generator=g)
X = (X / X.std()).to(dev)
torch.manual_seed(0)
m = V.GaussianMixtureModel(20, 2)
m.initialize(X.cpu())
m.to(dev)
m.update(X, 20)
torch.cuda.synchronize()
n0 = _lib.lib().vbmp_launch_count()
t0 = time.perf_counter()
m.update(X, 100)
e = float(m.ELBO_last)
dt = time.perf_counter() - t0
print(f"{dt * 10:.3f} ms per iteration (wall, 100 iterations), {(_lib.lib().vbmp_launch_count() - n0) / 100:.1f} library launches per iteration, ELBO {e:.4f}")
t0 = time.perf_counter()
m.update(X, 100)
dt_q = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"host enqueue only: {dt_q * 10:.3f} ms per iteration")
_lib.profile_begin(4096)
m.update(X, 20)
torch.cuda.synchronize()
prof = _lib.profile_end()
print({k: (len(v) // 20, round(sum(a.elapsed_time(b) for a, b in v) / 20, 4)) for k, v in prof.items()})
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as pr:
    m.update(X, 20)
    torch.cuda.synchronize()
print(pr.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60))
cp = cProfile.Profile()
cp.enable()
m.update(X, 100)
cp.disable()
torch.cuda.synchronize()
pstats.Stats(cp).sort_stats("tottime").print_stats(45)
