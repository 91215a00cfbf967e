"""How hard are the responsibilities along the bench's cfg2 trajectory?  Fraction of samples whose row of p is not exactly
one-hot in fp16 x 2^14 (i.e. whose low weight image b is non-zero), and of 16- / 32-sample chunks with no such sample (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pyvbmp_b200 as V
dev = torch.device("cuda:0")
N, K, D = bench.ROWS_PER_GPU, bench.K, bench.D
for sep, tag in ((3.0, "cfg2 recipe (mu ~ 3 N(0,I))"), (0.3, "overlapping variant (mu ~ 0.3 N(0,I))")):
    if sep == 3.0:
        X = bench.synth_rows(N, dev, 1234)
    else:
        g = torch.Generator(device=dev).manual_seed(4321)
        mu = sep * torch.randn(K, D, generator=g, device=dev)
        X = torch.empty(N, D, device=dev)
        for a in range(0, N, 1 << 20):
            X[a:a + (1 << 20)] = mu[torch.randint(K, (1 << 20,), generator=g, device=dev)] + torch.randn(1 << 20, D, generator=g, device=dev)
    torch.manual_seed(0)
    m = V.GaussianMixtureModel(K, D)
    m.to(dev)
    m.initialize(X[: 1 << 20])
    for it in range(1, 26):
        m.update(X, 1)
        if it in (1, 2, 3, 5, 6, 10, 15, 20, 25):
            p = m.p
            s = (p * 16384.0)
            lo = (s - s.half().float()) != 0                  # non-zero low piece
            soft_rows = lo.any(-1)
            f = float(soft_rows.float().mean())
            c16 = float((~soft_rows.view(-1, 16).any(-1)).float().mean())
            c32 = float((~soft_rows.view(-1, 32).any(-1)).float().mean())
            # per 128-component block as well (a CTA pair covers 256 = all of them here)
            blk = lo.view(N, K // 128, 128).any(-1)
            c32b = float((~blk.view(-1, 32, K // 128).any(1)).float().mean())
            print(f"{tag} it {it:2d}: rows with a non-zero low image {f:.4f}; chunks without one: 16 rows {c16:.4f}, 32 rows {c32:.4f}, 32 rows x 128 components {c32b:.4f}; ELBO {float(m.ELBO_last):.6e}")
    del X, m
    torch.cuda.empty_cache()
