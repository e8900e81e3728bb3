"""CPU baseline in the REFERENCE'S ALGORITHMIC SHAPE  --  TEST/BENCH INFRASTRUCTURE ONLY (see dsm_oracle.py).

Times what one `train!` iteration costs per leaf in /root/reference as written, issuing the
same LAPACK/BLAS routines Julia's stdlib issues (here through SciPy/OpenBLAS, all host threads):

  fit!                 fit.jl:98,105,109/112/129   2 x update_cholesky!  = 2 x [Gram from stored P, +noise, dpotrf,
                                                    2 triangular solves]            (gaussianprocess.jl:82-108)
  mll!                 optimize.jl:27-39            dot + logdet
  updategradients!     fit.jl:306-311               Gram from P, ldiv!(cK, -I) (dpotrs with n RHS) + dger,
  ∇mll! -> ∇mll(gp)    optimize.jl:49 ->            1 dense n^3 GEMM for the sigma trace and one per length scale
                       gaussianprocess.jl:185-190   (kernels.jl:93,157,161) -- and the whole thing runs TWICE.

The distance tensor P is built once outside the timed region (the reference stores it in the GP,
gaussianprocess.jl:57).  `optimised=True` times the algorithmically minimal CPU version instead (one
factorisation, dpotri, O(n^2) traces) so that the GPU speed-up is not inflated by the reference's redundant work.
"""
from __future__ import annotations

import time
from typing import Dict, List, Sequence, Tuple

import numpy as np
import scipy.linalg as sla

from . import dsm_oracle as orc


def _gram_from_P(k: orc.Kernel, P: np.ndarray) -> np.ndarray:
    return orc.kernelmatrix_from_P(k, P)


def reference_leaf_iteration(x: np.ndarray, y: np.ndarray, k: orc.Kernel, logNoise: float) -> Tuple[float, float]:
    """Returns (seconds, lml) for one leaf, reference shape."""
    n = x.shape[0]
    P = orc.getdistancematrix(k, x)            # stored in the GP: not timed
    noise = np.exp(2 * logNoise) + orc.EPS_JITTER
    t0 = time.perf_counter()
    for _ in range(2):                          # fit!: update_cholesky! twice per leaf
        F = _gram_from_P(k, P)
        F[np.diag_indices(n)] += noise
        L, _ = sla.lapack.dpotrf(F, lower=1, overwrite_a=1)
        z = sla.solve_triangular(L, y, lower=True, check_finite=False)
        alpha = sla.solve_triangular(L, z, lower=True, trans="T", check_finite=False)
    lml = -(float(y @ alpha) + 2.0 * np.sum(np.log(np.diag(L))) + orc.LOG2PI * n) / 2.0
    for _ in range(2):                          # updategradients!(spn) and again inside ∇mll(gp)
        K = _gram_from_P(k, P)
        W = -np.eye(n)
        W = sla.cho_solve((L, True), W, overwrite_b=True, check_finite=False)     # ldiv!(cK, -I)
        W = sla.blas.dger(1.0, alpha, alpha, a=W, overwrite_a=1)                  # BLAS.ger!
        _ = np.exp(2 * logNoise) * np.trace(W)
        if k.type in (orc.ISO_SE, orc.ARD_SE):
            Ks = k.std() * K
            _ = 0.5 * np.trace(W @ (2.0 * Ks))                                    # kernels.jl:93,157
            if k.type == orc.ISO_SE:
                _ = 0.5 * np.trace(W @ (Ks * (P / np.exp(k.logl[0]) ** 2)))       # :96-97
            else:
                ls = np.exp(k.logl) ** 2
                for d in range(k.logl.size):
                    _ = 0.5 * np.trace((W @ Ks) * (P[:, :, d] / ls[d]))           # :161 (one GEMM per d)
        else:
            _ = 0.5 * np.trace(W @ (-2.0 * K))                                    # :198
    return time.perf_counter() - t0, lml


def optimised_leaf_iteration(x: np.ndarray, y: np.ndarray, k: orc.Kernel, logNoise: float) -> Tuple[float, float]:
    n = x.shape[0]
    noise = np.exp(2 * logNoise) + orc.EPS_JITTER
    t0 = time.perf_counter()
    F = orc.kernelmatrix_chunked(k, x)
    K = F.copy()
    F[np.diag_indices(n)] += noise
    L, _ = sla.lapack.dpotrf(F, lower=1, overwrite_a=1)
    z = sla.solve_triangular(L, y, lower=True, check_finite=False)
    alpha = sla.solve_triangular(L, z, lower=True, trans="T", check_finite=False)
    lml = -(float(y @ alpha) + 2.0 * np.sum(np.log(np.diag(L))) + orc.LOG2PI * n) / 2.0
    Fi, _ = sla.lapack.dpotri(L, lower=1)
    tr = np.trace(Fi)
    _ = float(alpha @ alpha) - tr
    _ = float(alpha @ (K @ alpha)) - (2 * np.sum(np.tril(Fi, -1) * np.tril(K, -1)) + np.sum(np.diag(Fi) * np.diag(K)))
    return time.perf_counter() - t0, lml


def sample_model_time(x: np.ndarray, y: np.ndarray, leaves_obs: Sequence[np.ndarray], leaf_means: Sequence[float],
                      kernel: orc.Kernel, logNoise: float, budget_s: float = 20.0, optimised: bool = False) -> Dict:
    """Stratified sample of leaves (evenly spaced ranks in the size-sorted order, small ones first), timed until the
    budget is spent, extrapolated to the whole model by sum n^3.  Returns evals/s and the sample description."""
    order = np.argsort([len(o) for o in leaves_obs])
    L = len(order)
    picks: List[int] = []
    for frac in (0.5, 0.25, 0.75, 0.1, 0.9, 0.4, 0.6, 0.0, 1.0):
        i = int(order[min(L - 1, int(frac * (L - 1)))])
        if i not in picks:
            picks.append(i)
    fn = optimised_leaf_iteration if optimised else reference_leaf_iteration
    spent, cost_s, used = 0.0, 0.0, []
    for i in picks:
        obs = leaves_obs[i]
        n = len(obs)
        est = spent / max(cost_s, 1.0) * n ** 3 if cost_s > 0 else 0.0
        if used and spent + est > budget_s:
            continue
        t, _ = fn(x[obs], y[obs] - leaf_means[i], kernel, logNoise)
        spent += t; cost_s += float(n) ** 3; used.append(n)
    cost_all = float(np.sum([float(len(o)) ** 3 for o in leaves_obs]))
    total = spent * cost_all / cost_s
    return {"evals_per_s": 1.0 / total, "seconds_per_eval": total, "sample_sizes": used, "sample_seconds": spent,
            "extrapolation": "sum n^3 over all %d leaves / sum n^3 over the sample" % L}
