"""Shared helpers for the parity tests: golden-fixture loading and tolerant comparisons."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def tag(fix, t):
    """Sub-dict of a fixture under ``t/`` with the prefix stripped, as torch tensors."""
    pre = t + "/"
    return {k[len(pre):]: torch.as_tensor(v) for k, v in fix.items() if k.startswith(pre)}


def relerr(a, b):
    """max |a-b| / max(|b|_inf, tiny): matrix-level relative error (the 1e-4 parity metric)."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def assert_close(a, b, rtol, what=""):
    e = relerr(a, b)
    assert e <= rtol, f"{what}: rel err {e:.3e} > {rtol:.1e}"


def argmax_mismatch_report(p_ours, p_ref, logits_ref=None):
    """Count argmax mismatches; for each, report the reference's own top-2 margin."""
    a = torch.as_tensor(p_ours).argmax(-1).cpu()
    b = torch.as_tensor(p_ref).argmax(-1).cpu()
    bad = (a != b).nonzero().flatten().tolist() if a.ndim == 1 else (a != b).nonzero().tolist()
    margins = []
    src = torch.as_tensor(p_ref if logits_ref is None else logits_ref).cpu()
    for i in bad[:32]:
        row = src[i] if a.ndim == 1 else src[tuple(i)]
        top = row.topk(2).values
        margins.append(float(top[0] - top[1]))
    return len(bad), margins


def assert_maxabs(a, b, tol, what=""):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    e = float((a - b).abs().max()) if a.numel() else 0.0
    assert e <= tol, f"{what}: max abs err {e:.3e} > {tol:.1e}"
