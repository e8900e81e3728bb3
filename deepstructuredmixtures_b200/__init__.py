"""deepstructuredmixtures_b200 -- B200-native (sm_100a) drop-in for the data-parallel hot path of
DeepStructuredMixtures.jl: the batched per-expert GP computation over the leaves of the region graph.

All numerics run in `libdsmgp.so` (hand-written CUDA, include/dsmgp.h).  Importing this package loads the
library and fails loudly when it is missing; there is no CPU fallback.
"""
from . import _native
from ._native import DsmgpError, PosDefException

import os as _os

# The CPU reference arm of bench.py needs the host-side region-graph builder (structure.py, pure NumPy) and must not map
# the CUDA library into its process; every compute entry point still loads it on first use and fails without it.
if _os.environ.get("DSMGP_STRUCTURE_ONLY") != "1":
    _native.lib()   # fail at import time when libdsmgp.so has not been built

from .kernels import ArdLinear, ArdSE, IsoLinear, IsoSE, KernelFunction, kernelmatrix  # noqa: E402
from .linalg import chol_continue_, chol_delete_rows, potrf_  # noqa: E402
from .model import (DSMGP, GaussianProcess, LeafGP, Model, PoE, buildBCM, buildDSMGP, buildPoE, evaluate, fit_,  # noqa: E402
                    fit_naive_, gPoE, grad_mll, infer_, leftGP, mll, mll_nodes, params, predict, prediction, rBCM,
                    reset_weights_, rightGP, setparams_, stats, update_, update_cholesky_)
from .structure import ConstMean, GPNode, GPSplitNode, GPSumNode, getLeaves, getOverlap  # noqa: E402
from .train import ADAM, Descent, RMSProp, finetune_, train_, train_gp_  # noqa: E402

__all__ = [n for n in dir() if not n.startswith("_")]
