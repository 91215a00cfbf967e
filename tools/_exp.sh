for r in 65536 131072 262144; do
python - <<PY
import sys, torch, time
sys.path.insert(0, '.')
import pyvbmp_b200 as V
V.MixtureofLinearTransforms.PREDICT_ROWS = $r
dev = torch.device("cuda:0")
n, p, K, N = 32, 32, 64, 1 << 20
g = torch.Generator(device=dev).manual_seed(0)
torch.manual_seed(0)
m = V.MixtureofLinearTransforms(n, p, K).to(dev)
X = torch.randn(N, p, 1, generator=g, device=dev)
Wt = torch.randn(K, n, p, generator=g, device=dev) / p ** 0.5
z = torch.randint(K, (N,), generator=g, device=dev)
Y = (torch.einsum("nij,nj->ni", Wt[z], X[..., 0]) + 0.1 * torch.randn(N, n, generator=g, device=dev)).unsqueeze(-1)
m.raw_update(X, Y, iters=2)
for _ in range(2): m.predict(X)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(4): m.predict(X)
torch.cuda.synchronize(); print("PREDICT_ROWS", $r, f"{(time.perf_counter() - t0) / 4 * 1e3:.2f} ms per call")
PY
done
