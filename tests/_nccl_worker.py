"""Worker of tests/test_nccl_sharded.py (launched by torch.distributed.run, one rank per GPU): sample-sharded GMM and MoLT
VB-EM through the product path (kernels + ONE NCCL all-reduce per iteration), compared on rank 0 with the single-rank run
on the concatenated rows; replicas must agree bit for bit."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyvbmp_b200 as V                     # noqa: E402
from pyvbmp_b200 import sharding            # noqa: E402

GMM_KEYS = ("dist.mu", "dist.lambda_mu", "dist.invU.invU", "dist.invU.U", "dist.invU.nu", "dist.invU.logdet_invU", "pi.alpha")
MOLT_KEYS = ("W.mu", "W.invV", "W.V", "W.invU.invU", "W.invU.U", "W.invU.nu", "pi.alpha")


def get(o, path):
    for a in path.split("."):
        o = getattr(o, a)
    return o


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    msgs = []

    def check(cond, msg):
        nonlocal ok
        if not cond:
            ok = False
            msgs.append(msg)

    def replicas_equal(t, what):
        buf = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(buf, t.contiguous())
        check(all(torch.equal(buf[0], b) for b in buf[1:]), f"replicas differ: {what}")

    # ---- GMM, d = 64, K = 256 (tcgen05 kernels), N = 32768 + a ragged tail ----------------------------------------
    N, d, K, iters = 32768 + 1000, 64, 256, 3
    g = torch.Generator().manual_seed(21)
    mu_t = 0.8 * torch.randn(K, d, generator=g)
    X = (mu_t[torch.randint(K, (N,), generator=g)] + torch.randn(N, d, generator=g)).to(dev)
    lo, hi = sharding.shard_rows(N, rank, world)

    def new_gmm():
        torch.manual_seed(0)                                   # replicated init: same seed on every rank
        m = V.GaussianMixtureModel(K, d).to(dev)
        m.dist.mu = X[torch.randint(N, (K,), generator=torch.Generator().manual_seed(5)).to(dev)].clone()
        return m
    sharding.enable()
    m = new_gmm()
    elbo = []
    for _ in range(iters):
        m.update(X[lo:hi], 1)
        elbo.append(m.ELBO_last.clone())
    for k in GMM_KEYS:
        replicas_equal(get(m, k), k)
    replicas_equal(torch.stack(elbo), "ELBO trace")
    sharding.disable()
    if rank == 0:
        s = new_gmm()
        elbo1 = []
        for _ in range(iters):
            s.update(X, 1)
            elbo1.append(s.ELBO_last.clone())
        for a, b in zip(elbo, elbo1):
            check(abs(float(a) - float(b)) <= 1e-6 * abs(float(b)), f"GMM ELBO sharded {float(a)} vs single {float(b)}")
        # free-running for 3 iterations: the statistics differ by summation order (<= 1e-6 per step, gated below) and the
        # trajectory amplifies that while assignments still move (3.2e-5 on mu measured here), so
        # the BASELINE tolerance applies; the strict gate is the one-step comparison further down
        for k in GMM_KEYS:
            check(relerr(get(m, k), get(s, k)) <= 1e-4, f"GMM {k}: {relerr(get(m, k), get(s, k)):.2e}")
        check(bool((m.assignment() == s.assignment()[lo:hi]).float().mean() > 0.9999), "GMM assignments")
    # one step from identical parameters: <= 1e-6 relative vs single-rank on the concatenated rows
    sharding.enable()
    a = new_gmm()
    a.update(X[lo:hi], 1)
    sharding.disable()
    if rank == 0:
        b = new_gmm()
        b.update(X, 1)
        check(torch.equal(a.p, b.p[lo:hi]), "GMM step-1 responsibilities are bit-identical per sample")
        check(abs(float(a.ELBO_last) - float(b.ELBO_last)) <= 1e-6 * abs(float(b.ELBO_last)), "GMM step-1 ELBO")
        for k in ("dist.mu", "dist.lambda_mu", "dist.invU.invU", "dist.invU.nu", "pi.alpha"):
            check(relerr(get(a, k), get(b, k)) <= 1e-6, f"GMM step-1 {k}: {relerr(get(a, k), get(b, k)):.2e}")

    # ---- MoLT, n = p = 32, K = 64 ------------------------------------------------------------------------------------
    N, n, p, K = 20000, 32, 32, 64
    g = torch.Generator().manual_seed(22)
    Xm = torch.randn(N, p, generator=g)
    Wt = torch.randn(K, n, p, generator=g) / p ** 0.5
    z = torch.randint(K, (N,), generator=g)
    Ym = torch.einsum("nij,nj->ni", Wt[z], Xm) + 0.1 * torch.randn(N, n, generator=g)
    Xm, Ym = Xm.unsqueeze(-1).to(dev), Ym.unsqueeze(-1).to(dev)
    lo, hi = sharding.shard_rows(N, rank, world)

    def new_molt():
        torch.manual_seed(1)
        return V.MixtureofLinearTransforms(n, p, K).to(dev)
    sharding.enable()
    a = new_molt()
    a.raw_update(Xm[lo:hi], Ym[lo:hi], iters=2)
    for k in MOLT_KEYS:
        replicas_equal(get(a, k), "MoLT " + k)
    replicas_equal(a.ELBO_last.reshape(1), "MoLT ELBO")
    sharding.disable()
    if rank == 0:
        b = new_molt()
        b.raw_update(Xm, Ym, iters=2)
        check(abs(float(a.ELBO_last) - float(b.ELBO_last)) <= 2e-6 * abs(float(b.ELBO_last)), "MoLT ELBO vs single rank")
        for k in MOLT_KEYS:
            check(relerr(get(a, k), get(b, k)) <= 1e-4, f"MoLT {k}: {relerr(get(a, k), get(b, k)):.2e}")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    for msg in msgs:
        print(f"[rank {rank}] FAIL {msg}", flush=True)
    if rank == 0:
        print("NCCL_SHARDED_OK" if int(flag) == 1 else "NCCL_SHARDED_FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
