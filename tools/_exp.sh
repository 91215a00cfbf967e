timeout 300 python -m pytest tests/test_cuda_kernels.py -m gpu -q -x -k "estep_kernels or gram_kernels" 2>&1 | tail -5
python - <<'PY'
import torch, time, pyvbmp_b200 as V
dev='cuda:0'
for K in (512, 1024):
    torch.manual_seed(0)
    N, d = 1<<19, 64
    g = torch.Generator(device=dev).manual_seed(1)
    mu = 3*torch.randn(K, d, generator=g, device=dev)
    X = mu[torch.randint(K,(N,),generator=g,device=dev)] + torch.randn(N,d,generator=g,device=dev)
    m = V.GaussianMixtureModel(K, d).to(dev); m.dist.mu = X[:K].clone()
    for _ in range(3): m.update(X,1)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(5): m.update(X,1)
    torch.cuda.synchronize(); print(f"K={K}: {(time.perf_counter()-t0)/5*1e3:.2f} ms per iteration, ELBO {float(m.ELBO_last):.6e}")
PY
