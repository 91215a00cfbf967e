"""Where the first update(X_pinned_host, 20) call after the single-iteration e2e section loses 80 - 170 ms on some boxes
(bench.py e2e_iters20.ms_per_call): per-iteration device and host timestamps inside the call, allocator statistics around it
(dev tool).  Mirrors bench.py's sequence: resident steps, 22 single-iteration host calls, update(Xh, 2), then timed calls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pyvbmp_b200 as V
from pyvbmp_b200.mixture import Mixture
dev = torch.device("cuda:0")
N, K, D = bench.ROWS_PER_GPU, bench.K, bench.D
X = bench.synth_rows(N, dev, 1234)
torch.manual_seed(0)
m = V.GaussianMixtureModel(K, D)
m.to(dev)
m.initialize(X[: 1 << 20])
for _ in range(8):
    m.update(X, 1)
Xh = torch.empty(X.shape, dtype=X.dtype, pin_memory=True)
Xh.copy_(X)
for _ in range(6):
    m.update(Xh, 1)
torch.cuda.synchronize()
m.update(Xh, 2)
torch.cuda.synchronize()

marks = []
orig = Mixture.update_assignments
def traced(self, Xa):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append((time.perf_counter(), e))
    return orig(self, Xa)
Mixture.update_assignments = traced

def stats():
    s = torch.cuda.memory_stats(dev)
    return {k: s[k] for k in ("num_alloc_retries", "num_device_alloc", "num_device_free", "reserved_bytes.all.current", "allocated_bytes.all.current")}

for call in range(3):
    marks.clear()
    s0 = stats()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    m.update(Xh, 20)
    float(m.ELBO_last)
    b.record(); b.synchronize()
    t1 = time.perf_counter()
    s1 = stats()
    dev_ms = [round(a.elapsed_time(e), 1) for _, e in marks]
    host_ms = [round((t - t0) * 1e3, 1) for t, _ in marks]
    print(f"call {call}: {a.elapsed_time(b):.1f} ms on the device, {(t1 - t0) * 1e3:.1f} ms wall")
    print("  start of resident iterations 2..20, device ms:", dev_ms)
    print("  same, host enqueue ms:                        ", host_ms)
    print("  allocator:", {k: (s1[k] - s0[k] if 'bytes' not in k else (round(s0[k] / 2**30, 2), round(s1[k] / 2**30, 2))) for k in s0})
