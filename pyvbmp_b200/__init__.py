"""pyvbmp_b200 — B200-native (sm_100a) replacement for pyVBMP's conjugate VB-EM hot path.

Host side mirrors the reference's class interfaces (same names, constructor signatures, attributes);
the arithmetic runs in hand-written CUDA kernels behind the C ABI of ``libvbmp_b200.so``
(include/vbmp_b200.h).  There is no CPU fallback.
"""
from .wishart import Wishart
from .niw import NormalInverseWishart
from .mnw import MatrixNormalWishart
from .gamma import Gamma, DiagonalWishart
from .normal_gamma import NormalGamma
from .mng import MatrixNormalGamma
from .dirichlet import Dirichlet
from .mixture import Mixture, GaussianMixtureModel
from .molt import MixtureofLinearTransforms
from .mvn import MultivariateNormal_vector_format, Delta
from .hmm import HMM, ARHMM, ARHMM_prXY, ARHMM_prXRY
from .install import install, uninstall, installed_classes
from . import sharding
from . import ops          # registers torch.ops.vbmp.*
from ._lib import VbmpError, LIB_PATH

__all__ = ["Wishart", "NormalInverseWishart", "MatrixNormalWishart", "Gamma", "DiagonalWishart", "NormalGamma",
           "MatrixNormalGamma", "Dirichlet", "Mixture",
           "GaussianMixtureModel", "MixtureofLinearTransforms", "HMM", "ARHMM", "install", "uninstall",
           "VbmpError", "LIB_PATH"]
