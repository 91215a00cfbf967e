timeout 600 python -m pytest tests -m gpu -q -k "empty_and_tiny" 2>&1 | tail -25
python - <<'PY'
import torch, pyvbmp_b200 as V
dev='cuda:0'
for iso in (False, True):
    torch.manual_seed(0)
    m = V.GaussianMixtureModel(8, 16, isotropic=iso).to(dev)
    for N in (0, 1, 3):
        m.update(torch.randn(N, 16, device=dev), 1)
        print('iso' if iso else 'full', N, float(m.ELBO_last), m.p.shape, float(m.NA.sum()))
torch.manual_seed(0)
t = V.MixtureofLinearTransforms(4, 3, 8).to(dev)
for N in (0, 1, 3):
    t.raw_update(torch.randn(N, 3, 1, device=dev), torch.randn(N, 4, 1, device=dev), iters=1)
    print('molt', N, float(t.ELBO_last), t.p.shape)
PY
