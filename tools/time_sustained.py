"""Dev tool: is the bench's 20-step figure the sustained one?  cfg2 on one GPU: 120 EM iterations in blocks of 10 with the SM
clock / power read back per block, then the host-rows calls update(X_host, iters) for iters = 1, 5, 20."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyvbmp_b200 as V
from bench import synth_rows, D, K

dev = torch.device("cuda:0")
N = 4_194_304
X = synth_rows(N, dev, 1234)
torch.manual_seed(0)
m = V.GaussianMixtureModel(K, D).to(dev)
m.dist.mu = X[torch.randint(1 << 20, (K,)).to(dev)].clone()


def smi():
    q = "clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap"
    return subprocess.run(["nvidia-smi", "--id=0", f"--query-gpu={q}", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()


for _ in range(3):
    m.update(X, 1)
torch.cuda.synchronize()
for blk in range(12):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        m.update(X, 1)
    b.record()
    s = smi()                     # sampled while the block is still running on the device
    b.synchronize()
    print(f"block {blk}: {a.elapsed_time(b) / 10:.2f} ms/iter   [{s}]", flush=True)
Xh = torch.empty(X.shape, pin_memory=True)
Xh.copy_(X)
for iters in (1, 1, 5, 20, 20):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    m.update(Xh, iters)
    float(m.ELBO_last)
    b.record(); b.synchronize()
    print(f"update(X_host, {iters}): {a.elapsed_time(b):.1f} ms total, {a.elapsed_time(b) / iters:.2f} ms/iter  [{smi()}]", flush=True)
