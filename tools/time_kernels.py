"""Time the E-step / Gram C-ABI calls alone on the cfg2 shape (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyvbmp_b200 import _lib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from check_umma import make, dev

N = int(os.environ.get("TK_N", 1 << 20)); d0 = int(os.environ.get("TK_D0", 64)); d1 = int(os.environ.get("TK_D1", 0)); K = int(os.environ.get("TK_K", 256))
z, z0, z1, W, m, cst, Dp = make(N, d0, d1, K, spread=3.0)
xg = torch.zeros(1, dtype=torch.int32, device=dev)
what = os.environ.get("TK_WHAT", "eg")
def timeit(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n
out = torch.empty((N, 1, K), device=dev)
if "e" in what:
    t = timeit(lambda: _lib.estep(z0, z1, N, 1, xg, W, m, cst, 1, K, Dp, 1, out=out))
    print(f"estep mode1 N={N} K={K} D={d0+d1}: {t:.3f} ms  ({2.0*N*K*(d0+d1)**2/t/1e9:.1f} algorithmic TFLOP/s)")
    t = timeit(lambda: _lib.estep(z0, z1, N, 1, xg, W, m, cst, 1, K, Dp, 0, out=out))
    print(f"estep mode0 N={N} K={K} D={d0+d1}: {t:.3f} ms  (logits only)")
if "g" in what:
    p, lzn, NA, lZ = _lib.estep(z0, z1, N, 1, xg, W, m, cst, 1, K, Dp, 1)
    t = timeit(lambda: _lib.gram(z0, z1, N, 1, xg, p.view(N, 1, K), 1, xg, 1, K, Dp))
    print(f"gram N={N} K={K} D={d0+d1}: {t:.3f} ms  ({2.0*N*K*(d0+d1)**2/t/1e9:.1f} algorithmic TFLOP/s)  FL={os.environ.get('VBMP_GRAM_FL','')}")
