"""Drop the CUDA path in behind an imported pyVBMP tree, without editing it.

The reference binds its node classes into module globals at import time
(``import dists.NormalInverseWishart as NormalInverseWishart`` — models/GaussianMixtureModel.py:2-4,
transforms/MixtureofLinearTransforms.py:5-8, models/ARHMM.py:7-11), so install() walks the loaded
``dists.* / transforms.* / models.*`` modules and rebinds every global that *is* the reference
``NormalInverseWishart`` / ``Wishart`` / ``MatrixNormalWishart`` class to the replacement, and
patches the fused E-step onto the reference ``Mixture`` / ``MixtureofLinearTransforms`` classes
(SURVEY.md §1 "verified install mechanism").  The reference models then run unchanged on top of
libvbmp_b200.so; construct them under ``torch.set_default_device('cuda')``.
"""
from __future__ import annotations

import sys

_undo = []


def install(reference_root=None, fuse_assignments=True, verbose=False):
    """Returns the number of module globals rebound."""
    from .niw import NormalInverseWishart
    from .wishart import Wishart
    from .mnw import MatrixNormalWishart
    from . import mixture as _mix, molt as _molt

    if reference_root is not None and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    import dists       # noqa: F401  (the reference's top-level packages)
    import transforms  # noqa: F401
    import models      # noqa: F401

    ref = {
        "NormalInverseWishart": sys.modules["dists.NormalInverseWishart"].NormalInverseWishart
        if hasattr(sys.modules.get("dists.NormalInverseWishart"), "NormalInverseWishart") else dists.NormalInverseWishart,
        "Wishart": sys.modules["dists.Wishart"].Wishart if hasattr(sys.modules.get("dists.Wishart"), "Wishart")
        else dists.Wishart,
        "MatrixNormalWishart": sys.modules["transforms.MatrixNormalWishart"].MatrixNormalWishart
        if hasattr(sys.modules.get("transforms.MatrixNormalWishart"), "MatrixNormalWishart")
        else transforms.MatrixNormalWishart,
    }
    new = {"NormalInverseWishart": NormalInverseWishart, "Wishart": Wishart, "MatrixNormalWishart": MatrixNormalWishart}
    n = 0
    for name, mod in list(sys.modules.items()):
        if mod is None or not (name in ("dists", "transforms", "models")
                               or name.startswith(("dists.", "transforms.", "models."))):
            continue
        for attr, val in list(vars(mod).items()):
            for key, cls in ref.items():
                if val is cls:
                    setattr(mod, attr, new[key])
                    _undo.append((mod, attr, cls))
                    n += 1
                    if verbose:
                        print(f"rebound {name}.{attr} -> pyvbmp_b200.{key}")
    if fuse_assignments:
        Mixture = sys.modules["dists.Mixture"].Mixture if hasattr(sys.modules.get("dists.Mixture"), "Mixture") \
            else dists.Mixture
        MoLT = sys.modules["transforms.MixtureofLinearTransforms"].MixtureofLinearTransforms \
            if hasattr(sys.modules.get("transforms.MixtureofLinearTransforms"), "MixtureofLinearTransforms") \
            else transforms.MixtureofLinearTransforms
        _undo.append((Mixture, "update_assignments", Mixture.update_assignments))
        Mixture.update_assignments = _mix.fused_update_assignments
        _undo.append((MoLT, "update_assignments", MoLT.update_assignments))
        MoLT.update_assignments = _molt.fused_update_assignments
    return n


def uninstall():
    while _undo:
        obj, attr, val = _undo.pop()
        setattr(obj, attr, val)
