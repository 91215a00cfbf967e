"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the shape planner
flattens the reference's conventions correctly, constructors replicate the reference's RNG
consumption, and the product path refuses to run without CUDA (no silent CPU fallback)."""
import ctypes
import os
import re
import sys

import pytest
import torch

import pyvbmp_b200 as V
from pyvbmp_b200 import _lib, _shapes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "vbmp_b200.h")).read()
    declared = set(re.findall(r"\b(vbmp_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().vbmp_version() == 2


def test_plan_gmm_and_replicas():
    p = _shapes.make_plan((256,), (), (1,), (1000,))
    assert (p.K, p.G, p.GX, p.xg, p.pg, p.k_is_batch) == (256, 1, 1, (0,), (0,), True)
    # Mixture over NIW(batch=(3,6)), data (N,3,2) viewed (N,3,1,2): tests/test_dists.py:261-276
    p = _shapes.make_plan((3, 6), (), (3, 1), (200,))
    assert (p.K, p.G, p.GX, p.xg, p.pg) == (6, 3, 3, (0, 1, 2), (0, 1, 2))
    # batch of HMMs: NIW(batch=(3,K)), data shared over replicas AND states: tests/test_models.py:353-356
    p = _shapes.make_plan((3, 5), (), (1, 1), (7, 4))
    assert (p.N, p.K, p.G, p.GX, p.xg, p.pg) == (28, 5, 3, 1, (0, 0, 0), (0, 1, 2))
    # event_dim>1: NIW(event=(3,2), batch=(5,)), data (N,1,3,2)
    p = _shapes.make_plan((5,), (3,), (1,), (150,))
    assert (p.K, p.G, p.GX, p.xg, p.pg) == (5, 3, 3, (0, 1, 2), (0, 0, 0))
    # per-component data (no shared mixture axis)
    p = _shapes.make_plan((2,), (), (2,), (40,))
    assert (p.K, p.G, p.GX, p.xg, p.k_is_batch) == (1, 2, 2, (0, 1), False)


def test_theta_roundtrip_and_logits_layout():
    plan = _shapes.make_plan((3, 5), (4,), (3, 1), (7,))
    t = torch.arange(3 * 5 * 4 * 2).float().view(3, 5, 4, 2)
    flat = _shapes.theta_to_GK(t, plan, 2, 1, 1)
    assert flat.shape == (3 * 4 * 5, 2)
    back = _shapes.GK_to_theta(flat, plan, (2,))
    assert torch.equal(back, t)
    # component (b=1, e=2, k=3) must sit at group g = b*4+e, index k
    assert torch.equal(flat.view(12, 5, 2)[1 * 4 + 2, 3], t[1, 3, 2])
    out = torch.arange(7 * 12 * 5).float().view(7, 12, 5)
    ref = _shapes.logits_to_ref(out, plan)
    assert ref.shape == (7, 3, 5, 4)
    assert ref[2, 1, 3, 2] == out[2, 1 * 4 + 2, 3]


def test_constructors_consume_rng_like_reference_shapes():
    torch.manual_seed(0)
    m = V.GaussianMixtureModel(5, 2)
    assert m.dist.mu.shape == (5, 2) and m.dist.lambda_mu.shape == (5,)
    assert m.dist.invU.invU.stride() == (0, 2, 1)          # stride-0 expanded prior, as in the reference
    assert m.dist.invU.nu.shape == (5,) and float(m.dist.invU.nu[0]) == 4.0
    torch.manual_seed(0)
    mu = torch.randn(5, 2)
    assert torch.equal(m.dist.mu, mu)
    d = V.NormalInverseWishart(event_shape=(3, 2), batch_shape=(5,))
    assert d.lambda_mu.shape == (5, 1) and d.invU.nu.shape == (5, 3) and d.invU.U.shape == (5, 3, 2, 2)
    w = V.MatrixNormalWishart(event_shape=(4, 5), batch_shape=(3,), pad_X=True)
    assert w.mu.shape == (3, 4, 6) and w.invV.shape == (3, 6, 6) and w.invU.U.shape == (3, 4, 4)
    h = V.ARHMM(4, 2, 3)
    assert h.transition.alpha.shape == (4, 4) and h.initial.alpha.shape == (4,)


def test_no_cpu_fallback():
    m = V.GaussianMixtureModel(4, 3)
    X = torch.randn(32, 3)
    with pytest.raises(V.VbmpError):
        m.update(X, 1)
    with pytest.raises(V.VbmpError):
        m.dist.Elog_like(X.view(32, 1, 3))
    with pytest.raises(V.VbmpError):
        m.dist.KLqprior()


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_install_rebinds_reference_globals():
    sys.path.insert(0, "/root/reference")
    try:
        n = V.install()
        import models
        import transforms
        assert n >= 20
        torch.manual_seed(0)
        g = models.GaussianMixtureModel(5, 2)
        cls = V.installed_classes()
        assert type(g.dist) is cls["NormalInverseWishart"] and isinstance(g.dist, V.NormalInverseWishart)
        assert type(g.dist.invU) is cls["Wishart"] and isinstance(g.dist.invU, V.Wishart)
        assert type(g).update_assignments.__name__ == "mixture_update_assignments"
        a = models.ARHMM(3, 2, 2)
        assert type(a.obs_dist) is cls["MatrixNormalWishart"] and isinstance(a.obs_dist, V.MatrixNormalWishart)
        assert transforms.MixtureofLinearTransforms.update_assignments.__name__ == "molt_update_assignments"
    finally:
        V.uninstall()
        sys.path.remove("/root/reference")
    import dists
    assert dists.NormalInverseWishart is not V.NormalInverseWishart


def test_torch_custom_ops_are_registered_with_shape_inference():
    """torch.ops.vbmp.* (pyvbmp_b200/ops.py): registered with the dispatcher, fake-tensor shape functions in place."""
    import pyvbmp_b200.ops  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    for name in ("estep_logits", "estep_assign", "gram", "hmm_forward_backward", "rowgemm", "rowterm", "wsum", "moe_moments"):
        assert hasattr(torch.ops.vbmp, name)
    with FakeTensorMode():
        z, W, m, c = torch.empty(100, 8), torch.empty(5, 8, 8), torch.empty(5, 8), torch.empty(5)
        assert torch.ops.vbmp.estep_logits(z, None, W, m, c).shape == (100, 5)
        p, lzn, NA, lZ = torch.ops.vbmp.estep_assign(z, torch.empty(100, 4), torch.empty(5, 16, 16), torch.empty(5, 16), c)
        assert p.shape == (100, 5) and lzn.shape == (100,) and NA.shape == (5,) and lZ.shape == ()
        assert torch.ops.vbmp.gram(z, torch.empty(100, 3), p, False).shape == (5, 12, 12)
        out = torch.ops.vbmp.hmm_forward_backward(torch.empty(7, 3, 4), torch.empty(4, 4), torch.empty(4))
        assert [tuple(t.shape) for t in out] == [(7, 3, 4), (3, 4, 4), (3, 4), (3,)]
        assert torch.ops.vbmp.rowgemm(torch.empty(100, 9), torch.empty(9, 40), torch.empty(40)).shape == (100, 40)
        assert torch.ops.vbmp.rowterm(torch.empty(100, 64), torch.empty(64, 5), torch.empty(100, 5), -0.5).shape == (100, 5)
        assert torch.ops.vbmp.wsum(torch.empty(100, 8), torch.empty(100, 64)).shape == (8, 64)
        mu, Sig = torch.ops.vbmp.moe_moments(torch.empty(100, 8, 16), torch.empty(100, 8), None)
        assert mu.shape == (100, 16) and Sig.shape == (100, 16, 16)


def test_bench_clock_sampler_counts_only_the_timed_region():
    """bench.py starts nvidia-smi during warm-up (it needs a few hundred ms before its first line) and must report only the
    samples taken after mark(); with none inside a very short region it falls back to the last one before it and says so."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class FakeProc:
        def terminate(self): pass
        def wait(self, timeout=None): pass

    def row(mhz, cap):
        return ["0", str(mhz), "1965", "900.0", "0x4", "Not Active", "Not Active", "Not Active", "Active" if cap else "Not Active"]
    s = bench.ClockSampler(0)
    s.proc = FakeProc()
    s.rows = [row(1965, False), row(1900, False)]           # warm-up
    s.mark()
    s.rows += [row(1500, True), row(1520, True), row(1540, True)]
    out = s.stop()
    assert out["samples"] == 3 and out["sm_mhz"] == 1520.0 and out["reasons"] == ["sw_power_cap"] and "note" not in out
    s = bench.ClockSampler(0)
    s.proc = FakeProc()
    s.rows = [row(1965, False), row(1700, True)]
    s.mark()
    out = s.stop()
    assert out["samples"] == 1 and out["sm_mhz"] == 1700.0 and "note" in out
