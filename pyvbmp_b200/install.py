"""Drop the CUDA path in behind an imported pyVBMP tree, without editing it and without taking anything away.

The reference binds its node classes into module globals at import time
(``import dists.NormalInverseWishart as NormalInverseWishart`` — models/GaussianMixtureModel.py:2-4,
transforms/MixtureofLinearTransforms.py:5-8, models/ARHMM.py:7-11), so install() walks the loaded
``dists.* / transforms.* / models.*`` modules and rebinds every global that *is* the reference
``NormalInverseWishart`` / ``Wishart`` / ``MatrixNormalWishart`` / ``NormalGamma`` / ``MatrixNormalGamma`` class to an
INSTALLED class built here:

    class NormalInverseWishart(pyvbmp_b200.NormalInverseWishart, <reference NormalInverseWishart>)

* the hot methods (Elog_like / raw_update / ss_update / update / KLqprior ...) run in libvbmp_b200.so when the node
  lives on a CUDA device;
* everything this library does not implement — the message-passing methods of MatrixNormalWishart
  (transforms/MatrixNormalWishart.py:251-398), any future method — resolves to the reference class through the MRO;
* a node on the CPU keeps the reference's own code for the hot methods too (this library has no CPU path), and a
  MatrixNormalWishart constructed with ``mask`` / ``X_mask`` (:98-120) IS a plain reference object.

So LinearDynamicalSystems, DynamicMarkovBlanketDiscovery, dMixtureofLinearTransforms, HHMM ... construct and run after
install() exactly as before, while GaussianMixtureModel, MixtureofLinearTransforms, ARHMM and the HMM family pick up the
kernels.  The fused E-step (K1 + K2 with the softmax epilogue) is patched onto the reference ``Mixture`` /
``MixtureofLinearTransforms`` classes and the forward-backward kernel onto ``HMM``; each patch falls back to the
reference's own method for layouts / devices it does not take.  Construct models under
``torch.set_default_device('cuda')`` (the reference creates constants on the default device).
"""
from __future__ import annotations

import sys

_undo = []
_installed = {}


def _ref_class(modname, clsname, pkg):
    mod = sys.modules.get(modname)
    if mod is not None and hasattr(mod, clsname) and isinstance(getattr(mod, clsname), type):
        return getattr(mod, clsname)
    return getattr(pkg, clsname)        # the package attribute IS the class (the package __init__ shadows the submodule)


def _dispatch(ours, ref, name, probe):
    """Method ``name``: this library's implementation for CUDA-resident nodes, the reference's own otherwise."""
    f_ours, f_ref = getattr(ours, name), getattr(ref, name)

    def method(self, *a, **k):
        if probe(self).is_cuda:
            return f_ours(self, *a, **k)
        return f_ref(self, *a, **k)
    method.__name__ = name
    method.__qualname__ = f"{ours.__name__}.{name}"
    method.__doc__ = f_ours.__doc__
    return method


def _build_classes(refs):
    from .niw import NormalInverseWishart
    from .wishart import Wishart
    from .mnw import MatrixNormalWishart
    from .normal_gamma import NormalGamma
    from .mng import MatrixNormalGamma

    hot = {
        "Wishart": (Wishart, lambda s: s.invU, ("ss_update", "ElogdetinvSigma", "KLqprior")),
        "NormalInverseWishart": (NormalInverseWishart, lambda s: s.mu, ("ss_update", "raw_update", "Elog_like", "KLqprior")),
        "MatrixNormalWishart": (MatrixNormalWishart, lambda s: s.mu,
                                ("ss_update", "raw_update", "update", "Elog_like", "Elog_like_given_pX_pY", "KLqprior",
                                 "predict")),
        "NormalGamma": (NormalGamma, lambda s: s.mu, ("raw_update", "Elog_like")),
        "MatrixNormalGamma": (MatrixNormalGamma, lambda s: s.mu,
                              ("ss_update", "raw_update", "update", "Elog_like", "Elog_like_given_pX_pY", "KLqprior",
                               "predict")),
    }
    out = {}
    for key, (ours, probe, names) in hot.items():
        ref = refs[key]
        ns = {n: _dispatch(ours, ref, n, probe) for n in names if hasattr(ref, n)}
        ns["__doc__"] = f"pyvbmp_b200.{key} installed over the reference class (see pyvbmp_b200/install.py)."
        ns["__module__"] = ours.__module__
        if key in ("MatrixNormalWishart", "MatrixNormalGamma"):
            # transforms/MatrixNormalWishart.py:20: (event_shape, batch_shape, prior_parms, scale, mask, X_mask, ...);
            # transforms/MatrixNormalGamma.py:22-24: (..., scale, uniform_precision, mask, X_mask, ...)
            def _make_new(ref, at):
                def __new__(cls, *a, **k):
                    mask = k.get("mask", a[at] if len(a) > at else None)
                    X_mask = k.get("X_mask", a[at + 1] if len(a) > at + 1 else None)
                    if mask is not None or X_mask is not None:
                        return ref(*a, **k)         # masked nodes are outside the accelerated path: a reference object
                    return object.__new__(cls)
                return __new__
            ns["__new__"] = _make_new(ref, 4 if key == "MatrixNormalWishart" else 5)
        out[key] = type(key, (ours, ref), ns)
    return out


def install(reference_root=None, fuse_assignments=True, fuse_hmm=True, verbose=False):
    """Returns the number of module globals rebound."""
    from . import mixture as _mix, molt as _molt, hmm as _hmm

    if _undo:
        uninstall()
    if reference_root is not None and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    import dists       # noqa: F401  (the reference's top-level packages)
    import transforms  # noqa: F401
    import models      # noqa: F401

    refs = {
        "NormalInverseWishart": _ref_class("dists.NormalInverseWishart", "NormalInverseWishart", dists),
        "Wishart": _ref_class("dists.Wishart", "Wishart", dists),
        "MatrixNormalWishart": _ref_class("transforms.MatrixNormalWishart", "MatrixNormalWishart", transforms),
        "NormalGamma": _ref_class("dists.NormalGamma", "NormalGamma", dists),
        "MatrixNormalGamma": _ref_class("transforms.MatrixNormalGamma", "MatrixNormalGamma", transforms),
    }
    new = _build_classes(refs)
    _installed.clear()
    _installed.update(new)
    n = 0
    for name, mod in list(sys.modules.items()):
        if mod is None or not (name in ("dists", "transforms", "models")
                               or name.startswith(("dists.", "transforms.", "models."))):
            continue
        for attr, val in list(vars(mod).items()):
            for key, cls in refs.items():
                if val is cls:
                    setattr(mod, attr, new[key])
                    _undo.append((mod, attr, cls))
                    n += 1
                    if verbose:
                        print(f"rebound {name}.{attr} -> pyvbmp_b200.{key}")
    # nodes built inside this library's own constructors (NIW / MNW make their Wishart) must be installed classes too,
    # so that the reference's methods find the reference's interface on them
    from . import niw as _niw, mnw as _mnw
    for m in (_niw, _mnw):
        _undo.append((m, "Wishart", m.Wishart))
        m.Wishart = new["Wishart"]
    if fuse_assignments:
        Mixture = _ref_class("dists.Mixture", "Mixture", dists)
        MoLT = _ref_class("transforms.MixtureofLinearTransforms", "MixtureofLinearTransforms", transforms)
        orig_mix, orig_molt = Mixture.update_assignments, MoLT.update_assignments

        def mixture_update_assignments(self, X):
            return _mix.fused_update_assignments(self, X, fallback=orig_mix)

        def molt_update_assignments(self, X, Y):
            return _molt.fused_update_assignments(self, X, Y, fallback=orig_molt)
        _undo.append((Mixture, "update_assignments", orig_mix))
        Mixture.update_assignments = mixture_update_assignments
        _undo.append((MoLT, "update_assignments", orig_molt))
        MoLT.update_assignments = molt_update_assignments
    if fuse_hmm:
        HMM = _ref_class("models.HMM", "HMM", models)
        orig_fb = HMM.forward_backward_logits

        def forward_backward_logits(self, fw_logits):
            if fw_logits.is_cuda and fw_logits.shape[-1] <= 32 and fw_logits.numel() > 0:
                return _hmm.HMM.forward_backward_logits(self, fw_logits)
            return orig_fb(self, fw_logits)
        _undo.append((HMM, "forward_backward_logits", orig_fb))
        HMM.forward_backward_logits = forward_backward_logits
    return n


def installed_classes():
    """The classes the last install() bound in place of the reference's (empty before install / after uninstall)."""
    return dict(_installed)


def uninstall():
    while _undo:
        obj, attr, val = _undo.pop()
        setattr(obj, attr, val)
    _installed.clear()
