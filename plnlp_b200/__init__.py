"""plnlp_b200 -- B200-native (sm_100a) training / scoring hot path of PLNLP.

Same Python surface as the reference's ``plnlp`` package for the in-scope path (SURVEY.md
section 8): ``model.BaseModel``, ``layer.SAGE/GCN/MLPPredictor/DotPredictor``, ``loss.*``,
``negative_sample.*``, ``utils.*``, ``logger.Logger``.  All arithmetic runs in hand-written CUDA
kernels reached through the C ABI in ``include/plnlp_b200.h``; there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .graph import CSRGraph, SparseTensor  # noqa: F401

__version__ = "0.1.0"
