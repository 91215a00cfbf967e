"""2-rank NCCL run of the sample-sharded product path (SURVEY.md §8e, Appendix D "2/4/8-rank sharded vs single-rank"):
replicas bit-identical after every iteration, and within 1e-6 (one step) of the single-rank run on the concatenated
rows.  Needs two GPUs on the box: `gpurun --gpus 2 -- python -m pytest tests/test_nccl_sharded.py -m gpu`."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2])
def test_sharded_vbem_over_nccl_matches_single_rank(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(HERE, "_nccl_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "NCCL_SHARDED_OK" in out.stdout, (out.stdout[-3000:], out.stderr[-3000:])


@pytest.mark.gpu
def test_second_device_in_one_process():
    """A model whose tensors live on cuda:1 while the CUDA *current* device is cuda:0 (single-process multi-GPU use): every
    C-ABI call runs under a device guard and the library's caches are per device, so the results equal, bit for bit, those of
    the same model on cuda:0 — also when models on the two devices are updated alternately."""
    import pyvbmp_b200 as V
    if torch.cuda.device_count() < 2:
        pytest.skip(f"needs 2 GPUs, this box has {torch.cuda.device_count()}")
    torch.cuda.set_device(0)
    g = torch.Generator().manual_seed(21)
    K, d, N = 32, 64, 7000
    X = torch.randn(N, d, generator=g) * 1.2 + 0.3
    Xm, Ym = torch.randn(5000, 16, 1, generator=g), torch.randn(5000, 16, 1, generator=g)

    def gmm(dev):
        torch.manual_seed(3)
        m = V.GaussianMixtureModel(K, d)
        m.dist.mu = X[:K].clone()
        return m.to(dev)

    def molt(dev):
        torch.manual_seed(4)
        return V.MixtureofLinearTransforms(16, 16, 8).to(dev)

    a0, t0 = gmm("cuda:0"), molt("cuda:0")
    X0, Xm0, Ym0 = X.to("cuda:0"), Xm.to("cuda:0"), Ym.to("cuda:0")
    for _ in range(2):
        a0.update(X0, 1)
        t0.raw_update(Xm0, Ym0, iters=1)
    assert torch.cuda.current_device() == 0
    a1, t1 = gmm("cuda:1"), molt("cuda:1")
    X1, Xm1, Ym1 = X.to("cuda:1"), Xm.to("cuda:1"), Ym.to("cuda:1")
    b0 = gmm("cuda:0")                                         # a cuda:0 model updated between the cuda:1 calls
    for _ in range(2):
        a1.update(X1, 1)
        b0.update(X0, 1)
        t1.raw_update(Xm1, Ym1, iters=1)
    assert torch.cuda.current_device() == 0
    assert a1.p.device == torch.device("cuda:1") and a1.dist.mu.device == torch.device("cuda:1")
    for x, y in ((a0.p, a1.p), (a0.dist.mu, a1.dist.mu), (a0.dist.invU.invU, a1.dist.invU.invU), (a0.ELBO_last, a1.ELBO_last),
                 (a0.p, b0.p), (a0.ELBO_last, b0.ELBO_last), (t0.p, t1.p), (t0.W.mu, t1.W.mu), (t0.ELBO_last, t1.ELBO_last)):
        assert torch.equal(x.cpu(), y.cpu())
