"""CPU, world_size=2, gloo: the sample-sharded path's host logic.  Each rank owns a contiguous slice of
the rows, builds its local packed block [Gram | logZ | NA], ONE all-reduce sums it, and the replicated update must then be
identical on every rank and equal to the single-rank result.  The per-rank arithmetic here comes from the CPU oracle
(the CUDA kernels need a GPU); what is under test is pyvbmp_b200.sharding (row split, packing, collective, broadcast)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _local_block(X, mu, K):
    """Responsibilities from isotropic logits (a stand-in E-step) and the weighted Gram of [x;1] in fp64."""
    L = -0.5 * ((X[:, None, :] - mu[None]) ** 2).sum(-1)
    lz = torch.logsumexp(L, -1)
    p = (L - lz[:, None]).exp()
    Z1 = torch.cat([X, torch.ones(X.shape[0], 1, dtype=X.dtype)], -1)
    G = torch.einsum("nk,ni,nj->kij", p, Z1, Z1)
    return G, lz.sum().reshape(()), p.sum(0)


def _worker(rank, world, port, N, d, K, out):
    import sys
    sys.path.insert(0, ROOT)
    from pyvbmp_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sharding.enable()
        assert sharding.enabled()
        g = torch.Generator().manual_seed(11)
        X = torch.randn(N, d, generator=g, dtype=torch.float64) * 1.3 + 0.2
        # replicated init: every rank proposes its own means, rank 0's are broadcast once
        mu = torch.randn(K, d, generator=torch.Generator().manual_seed(100 + rank), dtype=torch.float64)
        sharding.broadcast_(mu, 0)
        lo, hi = sharding.shard_rows(N, rank, world)
        G, logZ, NA = _local_block(X[lo:hi], mu, K)
        G2, logZ2, NA2 = sharding.all_reduce_packed([G.float(), logZ.float(), NA.float()])
        torch.save({"mu": mu, "G": G2, "logZ": logZ2, "NA": NA2, "rows": (lo, hi)}, out + f".{rank}")
    finally:
        sharding.disable()
        dist.destroy_process_group()


def test_shard_rows_cover_everything_once():
    from pyvbmp_b200 import sharding
    for n, w in [(10, 3), (4194304, 8), (7, 8), (0, 2)]:
        spans = [sharding.shard_rows(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(180)
def test_two_rank_allreduce_matches_single_rank(tmp_path):
    N, d, K, world = 1001, 5, 7, 2
    out = str(tmp_path / "res")
    port = _free_port()
    mp.spawn(_worker, args=(world, port, N, d, K, out), nprocs=world, join=True)
    r = [torch.load(out + f".{i}") for i in range(world)]
    # replicas agree bit for bit after the collective, and the means are rank 0's
    for k in ("mu", "G", "logZ", "NA"):
        assert torch.equal(r[0][k], r[1][k]), k
    assert r[0]["rows"] == (0, 501) and r[1]["rows"] == (501, 1001)
    # and equal the single-rank statistics on the concatenated rows
    g = torch.Generator().manual_seed(11)
    X = torch.randn(N, d, generator=g, dtype=torch.float64) * 1.3 + 0.2
    G, logZ, NA = _local_block(X, r[0]["mu"], K)
    assert float((r[0]["G"].double() - G).abs().max() / G.abs().max()) < 1e-6
    assert abs(float(r[0]["logZ"]) - float(logZ)) < 1e-6 * abs(float(logZ))
    assert float((r[0]["NA"].double() - NA).abs().max() / NA.abs().max()) < 1e-6
