"""CPU oracle for the DSMGP per-expert GP hot path  --  TEST INFRASTRUCTURE ONLY.

This file restates, in NumPy/SciPy FP64, the arithmetic of the reference Julia package
(trappmartin/DeepStructuredMixtures, /root/reference/src/*.jl) for the path that
`libdsmgp.so` accelerates.  It is the *checker*: only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs may import it.  The product
package (`deepstructuredmixtures_b200/`) never imports anything from `oracle/`.

PARITY UNPINNED (by the reference): the reference ships no tests, golden vectors or fixtures
 and no Julia toolchain exists in this image, so the reference
cannot be executed.  What pins this file instead: (i) the line-by-line restatement, each function
citing the reference file:line it follows, (ii) an INDEPENDENT 50-digit evaluation of small models
straight from the Julia formulas (tests/golden/make_mp_golden.py: mpmath only, imports nothing
from here; vectors committed as tests/golden/mp_golden.json and checked by
tests/test_oracle.py::test_oracle_against_independent_mpmath_vectors -- LML, as-written and
mathematical gradients of all four kernels and a mixture, mll!, the down-pass plain and
finetune-weighted, update!, infer!, DSMGP / PoE / gPoE / rBCM predictions), (iii) finite-difference
and LAPACK self-checks in tests/test_oracle.py, and (iv) the restated in-source self tests
`test_chol_continue` and `lrtest` (AdvancedCholeskey.jl:61-135).

Third-party arithmetic the reference delegates to (not vendored, versions unpinned:
no Manifest.toml / [compat]):  LinearAlgebra stdlib -> LAPACK dpotrf/dtrtrs/dpotrs,
BLAS dger/dsyrk/dgemm (here: SciPy's LAPACK/BLAS);  Distances.jl pairwise(SqEuclidean())
(here: direct differences);  StatsFuns.logsumexp (here: max-shifted).

Conventions: all indices inside this module are 0-based; the C ABI / Julia side is
1-based (`leaf_obs`).  Matrices are NumPy row-major but every formula is layout-free.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import scipy.linalg as sla

EPS_JITTER = 1e-8          # DeepStructuredMixtures.jl:27   const ϵ = 1e-8
LOG2PI = math.log(2.0 * math.pi)

ISO_SE, ARD_SE, ISO_LINEAR, ARD_LINEAR = 0, 1, 2, 3
NODE_LEAF, NODE_SPLIT, NODE_SUM, NODE_KSUM = 0, 1, 2, 3


# ---------------------------------------------------------------------------------------
# kernels.jl
# ---------------------------------------------------------------------------------------
@dataclass
class Kernel:
    """IsoSE / ArdSE / IsoLinear / ArdLinear  (kernels.jl:59-64,109-114,174-177,209-212).

    `logl` is a length-1 (Iso) or length-D (Ard) array, `logs` = logσ (ignored for the
    linear kernels whose variance is fixed to 1, kernels.jl:181,216)."""
    type: int
    logl: np.ndarray
    logs: float = 0.0

    def __post_init__(self):
        self.logl = np.atleast_1d(np.asarray(self.logl, dtype=np.float64)).copy()
        self.logs = float(self.logs)

    @property
    def nparams(self) -> int:          # gaussianprocess.jl:139-145: (ℓ.., variance, noise)
        return self.logl.size + 2

    def variance(self) -> float:       # kernels.jl:68,118,181,216
        return math.exp(2.0 * self.logs) if self.type in (ISO_SE, ARD_SE) else 1.0

    def std(self) -> float:            # kernels.jl:69,119,182,217
        return math.exp(self.logs) if self.type in (ISO_SE, ARD_SE) else 1.0

    def copy(self) -> "Kernel":
        return Kernel(self.type, self.logl.copy(), self.logs)


def IsoSE(logl, logs):
    return Kernel(ISO_SE, [logl], logs)


def ArdSE(logl, logs):
    return Kernel(ARD_SE, logl, logs)


def IsoLinear(logl):
    return Kernel(ISO_LINEAR, [logl], 0.0)


def ArdLinear(logl):
    return Kernel(ARD_LINEAR, logl, 0.0)


def getdistancematrix(k: Kernel, x1: np.ndarray, x2: Optional[np.ndarray] = None) -> np.ndarray:
    """kernels.jl:55,83 (IsoSE: squared Euclidean), :137-144 (ArdSE: per-dim squared
    differences, n1 x n2 x D), :194 (IsoLinear: x1*x2'), :232 (ArdLinear: per-dim outer
    products; the reference returns a Vector of matrices)."""
    if x2 is None:
        x2 = x1
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    if k.type == ISO_SE:
        d = x1[:, None, :] - x2[None, :, :]
        return np.einsum("ijd,ijd->ij", d, d)
    if k.type == ARD_SE:
        d = x1[:, None, :] - x2[None, :, :]
        return d * d
    if k.type == ISO_LINEAR:
        return x1 @ x2.T
    if k.type == ARD_LINEAR:
        return x1[:, None, :] * x2[None, :, :]
    raise ValueError("unknown kernel type")


def kernelmatrix_from_P(k: Kernel, P: np.ndarray) -> np.ndarray:
    """kernels.jl:21-26 (Iso: K = v * kappa(P / l^2)),  :31-49 (Ard: K = v * sum_d kappa(P_d / l_d^2),
    accumulated in order d = 1..D -- ArdSE is ADDITIVE over dimensions, App. B Q1),
    rbfkernel :78  exp(-0.5*(z/l)),  linearkernel :189  z/l."""
    v = k.variance()
    if k.type in (ISO_SE, ISO_LINEAR):
        l = math.exp(k.logl[0]) ** 2
        K = np.exp(-0.5 * (P / l)) if k.type == ISO_SE else P / l
        return v * K
    ls = np.exp(k.logl) ** 2
    K = np.zeros(P.shape[:2])
    for d in range(P.shape[2]):
        K += np.exp(-0.5 * (P[:, :, d] / ls[d])) if k.type == ARD_SE else P[:, :, d] / ls[d]
    return v * K


def kernelmatrix(k: Kernel, x1: np.ndarray, x2: Optional[np.ndarray] = None) -> np.ndarray:
    """kernels.jl:15-18."""
    return kernelmatrix_from_P(k, getdistancematrix(k, x1, x2))


def kernelmatrix_chunked(k: Kernel, x1: np.ndarray, x2: Optional[np.ndarray] = None, chunk: int = 512) -> np.ndarray:
    """Same values as `kernelmatrix` without the n1*n2*D temporary (for oracle runs at n ~ 10^3)."""
    if x2 is None:
        x2 = x1
    out = np.empty((x1.shape[0], x2.shape[0]))
    for s in range(0, x1.shape[0], chunk):
        out[s:s + chunk] = kernelmatrix(k, x1[s:s + chunk], x2)
    return out


def kernel_gradients_as_written(k: Kernel, precomp: np.ndarray, K: np.ndarray, P: np.ndarray):
    """updategradients!(kernel, precomp, K, P) exactly as written, dense n^3 traces.

    IsoSE kernels.jl:85-99, ArdSE :146-164 (∂ℓ_d == 0 by operator precedence, App. B Q3),
    IsoLinear :196-200, ArdLinear :234-246 (reference is non-functional, App. B Q5; the
    definition used is SURVEY App. A.4).  Returns (dsigma, dl) like getgradients()."""
    if k.type == ISO_SE:
        s = k.std()
        l = math.exp(k.logl[0]) ** 2
        K = s * K                                   # lmul!(σ, K)                 :90
        ds = 0.5 * np.trace(precomp @ (2.0 * K))    # 0.5*tr(precomp * 2*K)       :93
        K = K * (P / l)                             # K .*= P/l                   :96
        dl = 0.5 * np.trace(precomp @ K)            #                             :97
        return ds, np.array([dl])
    if k.type == ARD_SE:
        s = k.std()
        ls = np.exp(k.logl) ** 2
        K = s * K                                   # :154
        ds = 0.5 * np.trace(precomp @ (2.0 * K))    # :157
        dl = np.zeros(k.logl.size)
        PK = precomp @ K
        for d in range(k.logl.size):
            # tr(precomp * K .* (p/ls[d]))  parses as  tr((precomp*K) .* (p/ls[d]))   :161
            dl[d] = 0.5 * np.trace(PK * (P[:, :, d] / ls[d]))
        return ds, dl
    if k.type == ISO_LINEAR:
        dl = 0.5 * np.trace(precomp @ (-2.0 * K))   # :198
        return 0.0, np.array([dl])
    if k.type == ARD_LINEAR:
        ls = np.exp(k.logl) ** 2                    # SURVEY A.4 definition (squared ℓ)
        dl = np.zeros(k.logl.size)
        for d in range(k.logl.size):
            Kd = P[:, :, d] / ls[d]
            dl[d] = 0.5 * np.trace(precomp @ (-2.0 * Kd))
        return 0.0, dl
    raise ValueError


def kernel_gradients_fast(k: Kernel, W: np.ndarray, K: np.ndarray, x: np.ndarray, mathematical: bool = False):
    """O(n^2) evaluation of the same quantities (tr(W K) = sum_ij W_ij K_ij for symmetric
    W, K).  With `mathematical=True` returns the true derivatives of the LML w.r.t. the
    log-parameters (no extra σ factor, ARD length-scale terms non-zero; App. B Q2/Q3)."""
    s = 1.0 if mathematical else k.std()
    if k.type == ISO_SE:
        l = math.exp(k.logl[0]) ** 2
        P = getdistancematrix(k, x)
        ds = s * np.sum(W * K)
        dl = 0.5 * s * np.sum(W * K * P) / l
        return ds, np.array([dl])
    if k.type == ARD_SE:
        ds = s * np.sum(W * K)
        dl = np.zeros(k.logl.size)
        if mathematical:
            v = k.variance()
            ls = np.exp(k.logl) ** 2
            for d in range(k.logl.size):
                Pd = (x[:, None, d] - x[None, :, d]) ** 2
                dl[d] = 0.5 * v * np.sum(W * np.exp(-0.5 * Pd / ls[d]) * Pd) / ls[d]
        return ds, dl
    if k.type == ISO_LINEAR:
        return 0.0, np.array([-np.sum(W * K)])
    if k.type == ARD_LINEAR:
        ls = np.exp(k.logl) ** 2
        dl = np.array([-np.sum(W * np.outer(x[:, d], x[:, d])) / ls[d] for d in range(k.logl.size)])
        return 0.0, dl
    raise ValueError


# ---------------------------------------------------------------------------------------
# gaussianprocess.jl
# ---------------------------------------------------------------------------------------
class GaussianProcess:
    """gaussianprocess.jl:14-80.  `y` is stored mean-subtracted (:72-74, means.jl:11-14)."""

    def __init__(self, x, y, mean: Optional[float] = None, kernel: Optional[Kernel] = None,
                 logNoise: float = math.log(7.0), run_cholesky: bool = False, centered: bool = False):
        self.x = np.asarray(x, dtype=np.float64).reshape(len(y), -1)
        y = np.asarray(y, dtype=np.float64)
        if centered:
            self.mean = float(mean)
            self.y = y.copy()
        else:
            self.mean = float(np.mean(y)) if mean is None else float(mean)   # :52 ConstMean(mean(y))
            self.y = y - self.mean
        self.kernel = kernel.copy() if kernel is not None else IsoSE(0.0, 0.0)   # :53
        self.logNoise = float(logNoise)
        self.N, self.D = self.x.shape
        self.L = None          # lower Cholesky factor (cK.L)
        self.alpha = None
        self.dnoise = 0.0      # ∂ϵ
        self.dsigma = 0.0
        self.dl = np.zeros(self.kernel.logl.size)
        if run_cholesky:
            self.update_cholesky()

    # -- parameters -----------------------------------------------------------------
    def noise(self) -> float:                              # :39  exp(2*logNoise)
        return math.exp(2.0 * self.logNoise)

    def nparams(self) -> int:                              # :139
        return self.kernel.nparams

    def params(self, logscale: bool = False):              # :141-145
        k = self.kernel
        if logscale:
            return k.logl.copy(), (k.logs if k.type in (ISO_SE, ARD_SE) else 0.0), self.logNoise
        return np.exp(k.logl), k.variance(), self.noise()

    def setparams(self, hyper: Sequence[float]):           # :153-161
        hyper = np.asarray(hyper, dtype=np.float64)
        assert hyper.size == self.nparams()
        self.logNoise = float(hyper[-1])
        if self.kernel.type in (ISO_SE, ARD_SE):           # setvariance! is a no-op for linear kernels :183,218
            self.kernel.logs = float(hyper[-2])
        self.kernel.logl[:] = hyper[:-2]

    # -- fit ------------------------------------------------------------------------
    def gram(self) -> np.ndarray:
        return kernelmatrix_chunked(self.kernel, self.x)

    def update_cholesky(self):
        """:82-108  F = K + (exp(2 logNoise) + 1e-8) I ; potrf!('L') ; α = L' \\ (L \\ y)."""
        F = self.gram()
        F[np.diag_indices_from(F)] += self.noise() + EPS_JITTER
        L, info = sla.lapack.dpotrf(F, lower=1, clean=1, overwrite_a=1)
        self.info = int(info)        # the reference discards info (:101)
        self.L = L
        z = sla.solve_triangular(L, self.y, lower=True, check_finite=False)
        self.alpha = sla.solve_triangular(L, z, lower=True, trans="T", check_finite=False)
        return self

    def mll(self) -> float:
        """:163  -(dot(y,α) + logdet(cK) + log2π*N)/2 ."""
        logdet = 2.0 * np.sum(np.log(np.diag(self.L)))
        return -(float(self.y @ self.alpha) + logdet + LOG2PI * self.N) / 2.0

    # -- gradients ------------------------------------------------------------------
    def W(self) -> np.ndarray:
        """ααinvcK! :219-226:  out = -I ; ldiv!(cK, out) ; ger!(1, α, α, out)  =>  αα' - F^{-1}."""
        if self.N > 3000:
            # same matrix through LAPACK dpotri (F^-1 from the factor, 2/3 n^3 instead of the 2 n^3 of potrs with n right-hand
            # sides): keeps the full-size parity tests (experts of 5,000-8,500 points) within seconds
            Fi, info = sla.lapack.dpotri(self.L, lower=1)
            assert info == 0
            Fi = np.tril(Fi) + np.tril(Fi, -1).T
            return np.outer(self.alpha, self.alpha) - Fi
        out = -np.eye(self.N)
        out = sla.cho_solve((self.L, True), out, check_finite=False)
        out += np.outer(self.alpha, self.alpha)
        return out

    def updategradients(self, as_written_dense: bool = False, mathematical: bool = False):
        """:165-178 + kernels.jl updategradients!.  `as_written_dense=True` follows the reference
        op-for-op (n^3 GEMM traces); otherwise the O(n^2) equivalent."""
        K = self.gram()
        W = self.W()
        self.dnoise = self.noise() * float(np.trace(W))          # :176
        if as_written_dense and not mathematical:
            P = getdistancematrix(self.kernel, self.x)
            self.dsigma, self.dl = kernel_gradients_as_written(self.kernel, W, K, P)
        else:
            self.dsigma, self.dl = kernel_gradients_fast(self.kernel, W, K, self.x, mathematical)
        return self

    def grad(self) -> np.ndarray:
        """∇mll! :206-217: grad = [∂ℓ..., ∂σ, ∂ϵ]."""
        return np.concatenate([np.atleast_1d(self.dl), [self.dsigma, self.dnoise]])

    def grad_mll(self, **kw) -> np.ndarray:                  # ∇mll(gp) :185-190
        self.updategradients(**kw)
        return self.grad()

    # -- prediction -----------------------------------------------------------------
    def prediction(self, xtest: np.ndarray, full_cov: bool = False):
        """:110-137.  μ = m + Knt'α ; V = L \\ Knt ; Σ = Ktt - V'V ; diag += exp(2 logNoise) (no 1e-8).
        Only diag(Σ) is consumed downstream (common.jl:136,147) so by default only it is formed."""
        xtest = np.asarray(xtest, dtype=np.float64).reshape(-1, self.D)
        Knt = kernelmatrix_chunked(self.kernel, self.x, xtest)
        mu = self.mean + Knt.T @ self.alpha
        V = sla.solve_triangular(self.L, Knt, lower=True, check_finite=False)
        if full_cov:
            S = kernelmatrix(self.kernel, xtest, xtest) - V.T @ V
            S[np.diag_indices_from(S)] += self.noise()
            return mu, S
        ktt = kernel_diag(self.kernel, xtest)
        return mu, ktt - np.sum(V * V, axis=0) + self.noise()


def kernel_diag(k: Kernel, x: np.ndarray) -> np.ndarray:
    """diag(kernelmatrix(k, x, x)) (gaussianprocess.jl:134 as consumed by common.jl:136)."""
    n = x.shape[0]
    if k.type == ISO_SE:
        return np.full(n, k.variance())
    if k.type == ARD_SE:
        return np.full(n, k.variance() * float(k.logl.size))      # v * Σ_d exp(0)
    if k.type == ISO_LINEAR:
        return np.sum(x * x, axis=1) / math.exp(k.logl[0]) ** 2
    ls = np.exp(k.logl) ** 2
    acc = np.zeros(n)
    for d in range(k.logl.size):
        acc += x[:, d] * x[:, d] / ls[d]
    return acc


# ---------------------------------------------------------------------------------------
# AdvancedCholeskey.jl
# ---------------------------------------------------------------------------------------
def gen_cov(D: int, rng: np.random.Generator) -> np.ndarray:
    """genCov :12  Symmetric(rand(D,D) .+ D*I, :L)."""
    A = rng.random((D, D)) + D * np.eye(D)
    return np.tril(A) + np.tril(A, -1).T


def chol_continue(A: np.ndarray, ki: int) -> Tuple[np.ndarray, int]:
    """chol_continue!(A, ki) :152-174, `ki` 1-based as in the reference: rows/cols 1..ki-1 hold a
    valid lower factor L11, the rest holds raw matrix entries.  L21 = A21 L11^{-T};
    A22 -= L21 L21' ; potrf!(A22) ; tril!(A).  Returns (A, info)."""
    A = np.tril(np.array(A, dtype=np.float64))
    k = ki - 1
    if k > 0:
        L11 = A[:k, :k]
        A[k:, :k] = sla.solve_triangular(L11, A[k:, :k].T, lower=True, check_finite=False).T
        A[k:, k:] -= np.tril(A[k:, :k] @ A[k:, :k].T)
    C, info = sla.lapack.dpotrf(A[k:, k:], lower=1, clean=1)
    if info == 0:
        A[k:, k:] = C
    return A, int(info)


def givens(f: float, g: float) -> Tuple[float, float, float]:
    """LinearAlgebra.givensAlgorithm for reals (LAPACK dlartg semantics): c*f + s*g = r, -s*f + c*g = 0."""
    if g == 0.0:
        return 1.0, 0.0, f
    if f == 0.0:
        return 0.0, 1.0, g
    r = math.hypot(f, g)
    if abs(f) > abs(g) and f < 0:
        r = -r
    return f / r, g / r, r


def lowrankupdate_as_written(A: np.ndarray, v: np.ndarray, k: int) -> np.ndarray:
    """lowrankupdate!(A, v, k, 'L') :20-59 verbatim, including the loop bound `for i = k:n` with
    n = length(v) (App. B Q7).  `k` is 1-based.  Kept to document the defect, not as a target."""
    A = A.copy(); v = v.copy()
    n = v.size
    assert A.shape[0] - (k - 1) == n
    for i in range(k, n + 1):              # 1-based i
        c, s, r = givens(A[i - 1, i - 1], v[i - k])
        A[i - 1, i - 1] = r
        for j in range(i + 1, n + 1):
            Aji = A[j - 1, i - 1]; vj = v[j - k]
            A[j - 1, i - 1] = c * Aji + s * vj
            v[j - k] = -s * Aji + c * vj
    return A


def chol_rank1_update_trailing(Lf: np.ndarray, v: np.ndarray, k: int) -> np.ndarray:
    """Corrected intent of lowrankupdate!: replace the trailing block L[k:, k:] (0-based k) by the
    lower Cholesky factor of  L[k:,k:] L[k:,k:]' + v v'   (Givens sweep over ALL columns k..d-1)."""
    Lf = Lf.copy(); v = np.array(v, dtype=np.float64)
    d = Lf.shape[0]
    assert v.size == d - k
    for i in range(k, d):
        c, s, r = givens(Lf[i, i], v[i - k])
        Lf[i, i] = r
        col = Lf[i + 1:, i].copy()
        vv = v[i - k + 1:].copy()
        Lf[i + 1:, i] = c * col + s * vv
        v[i - k + 1:] = -s * col + c * vv
    return Lf


def chol_delete_rows(Lf: np.ndarray, rows: Sequence[int]) -> np.ndarray:
    """Row/column deletion from a lower Cholesky factor (the operation fit.jl:179-195
    composes from lowrankupdate!).  `rows` 0-based.  Returns the factor of A[keep, keep]."""
    Lf = np.tril(np.array(Lf, dtype=np.float64))
    keep = [i for i in range(Lf.shape[0]) if i not in set(rows)]
    for i in sorted(rows):
        if i + 1 < Lf.shape[0]:
            Lf = chol_rank1_update_trailing(Lf, Lf[i + 1:, i].copy(), i + 1)
    return Lf[np.ix_(keep, keep)]


# ---------------------------------------------------------------------------------------
# Tree: node types (DeepStructuredMixtures.jl:40-71) rebuilt from the flat C-ABI description
# ---------------------------------------------------------------------------------------
@dataclass
class Node:
    type: int
    id: int
    children: List["Node"] = field(default_factory=list)
    logweights: Optional[np.ndarray] = None      # sum nodes
    split: Optional[List[Tuple[int, float]]] = None   # split nodes: [(d0based, s)...]
    gp: Optional[GaussianProcess] = None         # leaves
    obs: Optional[np.ndarray] = None             # leaves: 0-based ascending global rows
    kernelid: int = 0
    leaf_index: int = -1


def getLeaves(node: Node) -> List[Node]:
    """fit.jl:9-10 (depth-first, child order)."""
    if node.type == NODE_LEAF:
        return [node]
    out: List[Node] = []
    for c in node.children:
        out.extend(getLeaves(c))
    return out


def tree_from_flat(flat: dict, x: np.ndarray, y: np.ndarray, kernels: Sequence[Kernel], logNoise: float) -> Node:
    """Rebuild node objects from the flat arrays the C ABI consumes (include/dsmgp.h: dsmgp_tree):
    node_type, child_ptr, child_idx, leaf_of_node, split_dim, split_ptr, split_val, leaf_ptr,
    leaf_obs (1-based), leaf_kernel_id, leaf_mean, root.  `y` is the raw target vector."""
    nn = len(flat["node_type"])
    nodes: List[Optional[Node]] = [None] * nn
    for i in range(nn):       # children precede parents
        t = int(flat["node_type"][i])
        ch = [nodes[int(c)] for c in flat["child_idx"][flat["child_ptr"][i]:flat["child_ptr"][i + 1]]]
        nd = Node(type=t, id=i, children=ch)
        if t == NODE_LEAF:
            l = int(flat["leaf_of_node"][i])
            obs = np.asarray(flat["leaf_obs"][flat["leaf_ptr"][l]:flat["leaf_ptr"][l + 1]], dtype=np.int64) - 1
            kid = int(flat["leaf_kernel_id"][l])
            nd.obs, nd.kernelid, nd.leaf_index = obs, kid, l
            nd.gp = GaussianProcess(x[obs], y[obs], mean=float(flat["leaf_mean"][l]),
                                    kernel=kernels[kid], logNoise=logNoise)
        elif t == NODE_SPLIT:
            d = int(flat["split_dim"][i])
            sv = flat["split_val"][flat["split_ptr"][i]:flat["split_ptr"][i + 1]]
            nd.split = [(d, float(s)) for s in sv]
        else:
            nd.logweights = np.full(len(ch), -math.log(len(ch)))
        nodes[i] = nd
    return nodes[int(flat["root"])]


# ---------------------------------------------------------------------------------------
# optimize.jl : parameters, up-pass, down-pass
# ---------------------------------------------------------------------------------------
def setparams(node: Node, params: Sequence[float]):
    """optimize.jl:188-198.  A kernel-mixture sum slices θ per child kernel; every other inner node
    forwards the same θ to all children."""
    params = np.asarray(params, dtype=np.float64)
    if node.type == NODE_LEAF:
        node.gp.setparams(params)
    elif node.type == NODE_KSUM:
        c = 0
        for ch in node.children:
            n = ch.gp.nparams()
            ch.gp.setparams(params[c:c + n])
            c += n
    else:
        for ch in node.children:
            setparams(ch, params)


def logsumexp(v: Sequence[float]) -> float:
    """StatsFuns.logsumexp (max-shifted)."""
    v = np.asarray(v, dtype=np.float64)
    m = np.max(v)
    if not np.isfinite(m):
        return float(m)
    return float(m + math.log(np.sum(np.exp(v - m))))


def mll_up(node: Node, ell: dict) -> float:
    """mll!(node, ℓ) optimize.jl:27-39: leaf mll; split Σ children; sum logsumexp(-log K + child)
    with a UNIFORM prior (logweights ignored)."""
    if node.type == NODE_LEAF:
        v = node.gp.mll()
    elif node.type == NODE_SPLIT:
        v = 0.0
        for c in node.children:          # mapreduce(+) left to right
            v = v + mll_up(c, ell)
    else:
        K = len(node.children)
        v = logsumexp([-math.log(K) + mll_up(c, ell) for c in node.children])
    ell[node.id] = v
    return v


def grad_down(node: Node, dparent: float, lrho: float, ell: dict, logS: float, grad: np.ndarray,
              leaf_grads: dict, Drow: Optional[np.ndarray] = None):
    """∇mll!(node, ∇parent, lρ, ℓ, logS, ∇) optimize.jl:42-89 and the finetune variant :92-150
    (leaf contribution additionally scaled by D[g, leaf])."""
    if node.type == NODE_LEAF:
        w = math.exp(-logS + lrho + ell[node.id] + dparent)            # :48
        g = leaf_grads[node.leaf_index]
        if Drow is not None:
            grad += g * w * Drow[node.leaf_index]                       # :101
        else:
            grad += g * w                                               # :49
    elif node.type == NODE_SPLIT:
        for c in node.children:
            lp = ell[node.id] - ell[c.id]                               # :59
            grad_down(c, dparent + lp, lrho, ell, logS, grad, leaf_grads, Drow)
    elif node.type == NODE_SUM:
        K = len(node.children)
        for c in node.children:                                         # :72
            grad_down(c, -math.log(K) + dparent, math.log(K) + lrho, ell, logS, grad, leaf_grads, Drow)
    else:                                                               # kernel mixture :76-89
        c0 = 0
        for ch in node.children:
            n = ch.gp.nparams()
            grad_down(ch, dparent, lrho, ell, logS, grad[c0:c0 + n], leaf_grads, Drow)
            c0 += n


def nparams_total(root: Node) -> int:
    """n in train! optimisers.jl:15-17 (Σ over the left-most GP or the left-most kernel mixture)."""
    n = root
    while n.type in (NODE_SPLIT, NODE_SUM):
        n = n.children[0]
    if n.type == NODE_KSUM:
        return sum(c.gp.nparams() for c in n.children)
    return n.gp.nparams()


def fit(root: Node):
    """fit!(spn, D, gpmap) fit.jl:71-122 reduced to its observable result: every leaf ends with the
    exact factor of update_cholesky! (:105 always runs; the sharing branches either copy an identical
    factor or are numerically wrong, App. B Q6/Q7, so the exact factor is the parity target)."""
    for leaf in getLeaves(root):
        leaf.gp.update_cholesky()


def evaluate(root: Node, theta: Sequence[float], Drow: Optional[np.ndarray] = None,
             mathematical: bool = False, as_written_dense: bool = False):
    """One LML+gradient evaluation = optimisers.jl:43-77 without the Flux step:
    setparams! -> fit! -> mll! -> updategradients! -> ∇mll!.  Returns (lml_root, grad, ell, leaf_rows)
    where leaf_rows[l] = [lml_l, g_l...]."""
    setparams(root, theta)
    fit(root)
    ell: dict = {}
    mll_up(root, ell)
    leaves = getLeaves(root)
    leaf_grads = {lf.leaf_index: lf.gp.grad_mll(mathematical=mathematical, as_written_dense=as_written_dense)
                  for lf in leaves}
    grad = np.zeros(nparams_total(root))
    grad_down(root, 0.0, 0.0, ell, ell[root.id], grad, leaf_grads, Drow)
    rows = {lf.leaf_index: np.concatenate([[ell[lf.id]], leaf_grads[lf.leaf_index]]) for lf in leaves}
    return ell[root.id], grad, ell, rows


# ---------------------------------------------------------------------------------------
# fit.jl:12-39  overlap matrix
# ---------------------------------------------------------------------------------------
def getOverlap(root: Node, N: int) -> np.ndarray:
    """D[n,m] = 1 - |obs_n \\ obs_m| / |obs_n| for leaves under different children of a common sum
    node; the difference count is multiplied by (kernelid_n == kernelid_m), so D = 1 when the kernel
    ids differ (App. B Q15).  Leaf numbering = getLeaves order (gpmap, treeStructure.jl:421-426)."""
    leaves = getLeaves(root)
    L = len(leaves)
    D = np.zeros((L, L))
    bits = {}
    for lf in leaves:
        b = np.zeros(N, dtype=bool); b[lf.obs] = True
        bits[lf.leaf_index] = b

    def rec(node: Node) -> List[Node]:
        if node.type == NODE_LEAF:
            return [node]
        if node.type == NODE_SPLIT:
            out: List[Node] = []
            for c in node.children:
                out.extend(rec(c))
            return out
        r = [rec(c) for c in node.children]
        for i in range(len(r)):
            for j in range(i + 1, len(r)):
                for nn in r[i]:
                    for mm in r[j]:
                        bn, bm = bits[nn.leaf_index], bits[mm.leaf_index]
                        delta = np.logical_xor(bn, bm)
                        same = 1 if nn.kernelid == mm.kernelid else 0
                        dn = np.sum(delta & bn) * same
                        dm = np.sum(delta & bm) * same
                        D[nn.leaf_index, mm.leaf_index] = 1.0 - dn / np.sum(bn)
                        D[mm.leaf_index, nn.leaf_index] = 1.0 - dm / np.sum(bm)
        return [n for sub in r for n in sub]

    rec(root)
    return D


# ---------------------------------------------------------------------------------------
# common.jl : posterior weights and prediction
# ---------------------------------------------------------------------------------------
def update_weights(node: Node) -> float:
    """update!(node) common.jl:323-332.  Writes node.logweights on every sum node, returns z."""
    if node.type == NODE_LEAF:
        return node.gp.mll()
    if node.type == NODE_SPLIT:
        v = 0.0
        for c in node.children:
            v = v + update_weights(c)
        return v
    K = len(node.children)
    lw = np.array([-math.log(K) + update_weights(c) for c in node.children])
    z = logsumexp(lw)
    node.logweights = lw - z
    return z


def infer_weights(node: Node) -> float:
    """infer!(node) common.jl:336-355.  Kernel-mixture sum nodes (GPSumNode{T,GPNode}, :339-345) keep normalised posterior
    weights; sum nodes over sub-trees (:347-353) are reset to the uniform -log K after z is computed.  Returns z."""
    if node.type == NODE_LEAF:
        return node.gp.mll()
    if node.type == NODE_SPLIT:
        v = 0.0
        for c in node.children:
            v = v + infer_weights(c)
        return v
    K = len(node.children)
    lw = np.array([-math.log(K) + infer_weights(c) for c in node.children])
    z = logsumexp(lw)
    node.logweights = lw - z if node.type == NODE_KSUM else np.full(K, -math.log(K))
    return z


def reset_weights(node: Node) -> None:
    """reset_weights!(spn) common.jl:357-363: every sum node's logweights = -log K."""
    if node.type >= NODE_SUM:
        node.logweights = np.full(len(node.children), -math.log(len(node.children)))
    for c in node.children:
        reset_weights(c)


def fit_plan(root: Node, D: np.ndarray, tau: float = 0.05) -> List[Tuple[str, int]]:
    """The scheduling and case split of fit!(spn, D, gpmap; τ) fit.jl:71-122 and fitcontained! :124-292, WITHOUT the
    arithmetic: for every leaf (getLeaves order) the branch the reference takes and its main leaf:
        'main'      factored as somebody's main expert (:97-100) before being visited itself
        'self'      its own main / kernel ids differ / first(obs) < first(main.obs)      -> update_cholesky! (:107-112)
        'copy'      (true, true)    :132-143
        'delete'    (false, true)   :145-206 with length(toupdate)/nobs < τ  (else 'full')
        'continue'  (true, false)   :208-292 with the prefix test :246-248 and the τ test (else 'full')
        'full'      any branch that ends in update_cholesky!(jGP)"""
    leaves = getLeaves(root)
    n = len(leaves)
    S = [0] * n
    counts = [0] * n
    for j in range(n):
        i = int(np.argmax(D[:, j] * D[j, :]))                      # :79
        counts[i] += 1
        S[j] = i
    order = sorted(range(n), key=lambda l: counts[l])              # :86 (stable)
    processed = [False] * n
    out: List[Tuple[str, int]] = [("", -1)] * n
    for j in order:
        if processed[j]:
            continue
        i = S[j]
        if not processed[i]:
            processed[i] = True
            if i != j:
                out[i] = ("main", i)
        processed[j] = True
        jn, mn = leaves[j], leaves[i]
        jo, mo = jn.obs, mn.obs
        if i == j or mn.kernelid != jn.kernelid or jo[0] < mo[0]:  # :107-112
            out[j] = ("self", i)
            continue
        ione, jone = D[i, j] == 1.0, D[j, i] == 1.0
        if ione and jone:
            out[j] = ("copy", i)
        elif (not ione) and jone:
            minJ, maxJ, minM, maxM = jo[0], jo[-1], mo[0], mo[-1]
            assert minJ >= minM and maxJ <= maxM                    # :162-163
            e = len(mo) if maxJ == maxM else int(np.nonzero(mo == maxJ)[0][0]) + 1
            toupdate = np.setdiff1d(mo[:e], jo)
            out[j] = ("delete", i) if len(toupdate) / len(jo) < tau else ("full", i)
        elif ione and (not jone):
            minJ, maxJ, minM, maxM = jo[0], jo[-1], mo[0], mo[-1]
            assert minJ >= minM and maxJ >= maxM                    # :232-233
            s = 0 if minJ == minM else int(np.nonzero(mo == minJ)[0][0])
            k1 = int(np.nonzero(jo == maxM)[0][0]) + 1
            s1, s2 = jo[:k1], mo[s:]
            toupdate = np.setdiff1d(mo, s1)
            if len(s1) != len(s2) and minJ == minM:
                out[j] = ("full", i)
            else:
                out[j] = ("continue", i) if len(toupdate) / len(jo) < tau else ("full", i)
        else:
            out[j] = ("full", i)
    return out


def getchild(node: Node, x: np.ndarray) -> np.ndarray:
    """common.jl:101-122: child k (0-based here) iff s_{k-1} < x_d <= s_k ; first child iff x_d <= s_1.
    (A point above the last threshold makes the reference loop run out of bounds; callers never do that
    for the root region whose last threshold is +Inf.)"""
    T = x.shape[0]
    idx = np.full(T, -1, dtype=np.int64)
    for n in range(T):
        k = 0
        while idx[n] < 0:
            d, s = node.split[k]
            ok = x[n, d] <= s if k == 0 else (x[n, d] <= s) and (x[n, d] > node.split[k - 1][1])
            if ok:
                idx[n] = k
            k += 1
    return idx


def lse_rows(M: np.ndarray) -> np.ndarray:
    """lse(x; dims=2) common.jl:309-313."""
    m = np.max(M, axis=1, keepdims=True)
    return (np.log(np.sum(np.exp(M - m), axis=1, keepdims=True)) + m)[:, 0]


def _minpredict(node: Node, x: np.ndarray) -> np.ndarray:
    """common.jl:151-173."""
    if node.type == NODE_LEAF:
        return node.gp.prediction(x)[0]
    if node.type == NODE_SPLIT:
        idx = getchild(node, x)
        mu = np.zeros(x.shape[0])
        for k, c in enumerate(node.children):
            j = np.nonzero(idx == k)[0]
            mu[j] = _minpredict(c, x[j])
        return mu
    mu = np.full(x.shape[0], np.inf)
    for c in node.children:
        mu = np.minimum(mu, _minpredict(c, x))
    return mu


def _predict(node: Node, x: np.ndarray, mumin: np.ndarray):
    """common.jl:134-143 (leaf), :181-196 (split), :275-292 (sum)."""
    T = x.shape[0]
    if node.type == NODE_LEAF:
        mu, s2 = node.gp.prediction(x)
        s2 = np.where(s2 <= 0, EPS_JITTER, s2)                     # :137
        assert np.all(mu >= mumin)                                   # :138
        return np.log(mu - mumin), np.log(mu ** 2), np.log(s2)
    if node.type == NODE_SPLIT:
        idx = getchild(node, x)
        lm = np.zeros(T); lm2 = np.zeros(T); ls = np.zeros(T)
        for k, c in enumerate(node.children):
            j = np.nonzero(idx == k)[0]
            a, b, cc = _predict(c, x[j], mumin[j])
            lm[j], lm2[j], ls[j] = a, b, cc
        return lm, lm2, ls
    K = len(node.children)
    lm = np.zeros((T, K)); lm2 = np.zeros((T, K)); ls = np.zeros((T, K))
    for k, c in enumerate(node.children):
        a, b, cc = _predict(c, x, mumin)
        lm[:, k] = a + node.logweights[k]
        lm2[:, k] = b + node.logweights[k]
        ls[:, k] = cc + node.logweights[k]
    return lse_rows(lm), lse_rows(lm2), lse_rows(ls)


def predict_dsmgp(root: Node, x: np.ndarray):
    """predict(model::DSMGP, x) common.jl:294-304 (root sum) / :243-254 (root split) / :175-179 (leaf)."""
    x = np.asarray(x, dtype=np.float64)
    if root.type == NODE_SPLIT:
        idx = getchild(root, x)
        mu = np.zeros(x.shape[0]); s2 = np.zeros(x.shape[0])
        for k, c in enumerate(root.children):
            j = np.nonzero(idx == k)[0]
            mu[j], s2[j] = predict_dsmgp(c, x[j])
        return mu, s2
    mumin = _minpredict(root, x)
    lm, lm2, ls = _predict(root, x, mumin - 1.0)
    mu = np.exp(lm) + mumin - 1.0
    if root.type == NODE_LEAF:
        return mu, np.exp(ls)                                      # :178
    return mu, np.exp(ls) + (np.exp(lm2) - mu ** 2)                 # :299-300


def _predictPoE(node: Node, x: np.ndarray):
    """common.jl:145-149 (leaf: μ, 1/σ², no clamp), :198-208 (split: precision-weighted)."""
    if node.type == NODE_LEAF:
        mu, s2 = node.gp.prediction(x)
        return mu, 1.0 / s2
    if node.type != NODE_SPLIT:
        raise TypeError("MethodError: no method matching _predictPoE(::GPSumNode, ...)")
    mu = np.zeros(x.shape[0]); t = np.zeros(x.shape[0])
    for c in node.children:
        m_, t_ = _predictPoE(c, x)
        t += t_
        mu += t_ * m_
    return mu / t, t


def predict_poe(root: Node, x: np.ndarray):
    """predictPoE common.jl:256-260."""
    mu, t = _predictPoE(root, np.asarray(x, dtype=np.float64))
    return mu, 1.0 / t


def predict_gpoe(root: Node, x: np.ndarray):
    """predictgPoE / _predictgPoE common.jl:211-222,263-267: β = 1/#root-children on the root's children
    only (deeper levels use the plain PoE rule)."""
    x = np.asarray(x, dtype=np.float64)
    beta = 1.0 / len(root.children)
    mu = np.zeros(x.shape[0]); t = np.zeros(x.shape[0])
    for c in root.children:
        m_, t_ = _predictPoE(c, x)
        t += beta * t_
        mu += beta * t_ * m_
    mu = mu / t
    return mu, 1.0 / t


def leftGP(node: Node) -> GaussianProcess:
    """common.jl:124-127."""
    while node.type != NODE_LEAF:
        node = node.children[0]
    return node.gp


def predict_rbcm(root: Node, x: np.ndarray):
    """predictrBCM / _predictrBCM common.jl:224-241,269-273."""
    x = np.asarray(x, dtype=np.float64)
    gp = leftGP(root)
    s = kernel_diag(gp.kernel, x) + gp.noise()
    C = 1.0 / s
    mu = np.zeros(x.shape[0])
    for c in root.children:
        m_, t_ = _predictPoE(c, x)
        s_ = 1.0 / t_
        beta = 0.5 * (np.log(s) - np.log(s_))
        C = C + (beta * t_) - (beta / s)
        mu = mu + m_ * (beta * t_)
    mu = mu / C
    return mu, 1.0 / C
