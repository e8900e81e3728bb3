"""Public host API: a Python mirror of the reference's Julia interface for the hot path.

Julia is not available in this image, so the host side above the C ABI is written in Python with the
reference's names and argument meaning (a `!` suffix becomes a trailing underscore):

    buildDSMGP / buildPoE / buildBCM     treeStructure.jl:328-403
    GaussianProcess, update_cholesky_    gaussianprocess.jl:50-108
    fit_, fit_naive_                     fit.jl:67-122, 294-304
    mll, grad_mll (∇mll)                 optimize.jl:18-39, gaussianprocess.jl:163,185-217
    update_                              common.jl:323-334
    predict, prediction                  common.jl:294-307, gaussianprocess.jl:110-137
    params, setparams_, leftGP, rightGP  gaussianprocess.jl:139-161, optimize.jl:185-198, common.jl:124-132
    train_, finetune_                    optimisers.jl:4-145, finetuning.jl:3-88  (see train.py)

Every numerical step is a call into libdsmgp.so; nothing here computes a Gram matrix, factorisation,
gradient or prediction on the CPU.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _native as nat
from ._handle import Handle
from .kernels import ArdLinear, ArdSE, IsoLinear, IsoSE, KernelFunction
from .structure import (ConstMean, DSMGPConfig, GPNode, GPSplitNode, GPSumNode, Node, buildTree, flatten, getLeaves,
                        getOverlap)


class LeafGP:
    """View of one leaf expert (`node.dist` in the reference).  Parameters are read from / written to the
    device handle; `alpha`, `factors`, `mll` are fetched on demand."""

    def __init__(self, model: "Model", leaf: int):
        self.model, self.leaf = model, leaf
        self.node: GPNode = model.leaves[leaf]

    @property
    def kernel(self) -> KernelFunction:
        return self.node.kernel

    @property
    def N(self) -> int:
        return self.node.nobs

    def nparams(self) -> Tuple[int, int, int]:               # gaussianprocess.jl:139
        return (self.kernel.logl.size, 1, 1)

    def params(self, logscale: bool = False):                 # gaussianprocess.jl:141-145
        th = self.model.handle.get_leaf_params(self.leaf)
        k = self.kernel
        nl = k.logl.size
        logl, logs, logn = th[:nl], th[nl], th[nl + 1]
        iso = k.type in (nat.ISO_SE, nat.ISO_LINEAR)
        linear = k.type in (nat.ISO_LINEAR, nat.ARD_LINEAR)
        if logscale:
            return (float(logl[0]) if iso else logl.copy(), 0.0 if linear else float(logs), float(logn))
        ell = np.exp(logl)
        return (float(ell[0]) if iso else ell, 1.0 if linear else math.exp(2 * logs), math.exp(2 * logn))

    def setparams_(self, hyper: Sequence[float]):             # gaussianprocess.jl:153-161
        hyper = np.asarray(hyper, dtype=np.float64)
        self.model.handle.set_leaf_params(self.leaf, hyper)
        nl = self.kernel.logl.size
        self.kernel.logl[:] = hyper[:nl]
        if self.kernel.type in (nat.ISO_SE, nat.ARD_SE):
            self.kernel.logs = float(hyper[nl])
        self.node.logNoise = float(hyper[nl + 1])

    @property
    def alpha(self) -> np.ndarray:
        return self.model.handle.leaf_alpha(self.leaf)

    @property
    def factors(self) -> np.ndarray:
        return self.model.handle.leaf_factor(self.leaf)

    def mll(self) -> float:                                   # gaussianprocess.jl:163
        return float(self.model.handle.leaf_rows()[self.leaf, 0])

    def prediction(self, xtest) -> Tuple[np.ndarray, np.ndarray]:
        """prediction(gp, xtest) gaussianprocess.jl:131-137: returns (mu, diag(Sigma)).  The reference returns
        the full T x T Sigma but only its diagonal is ever consumed (common.jl:136,147)."""
        return self.model.handle.leaf_predict(self.leaf, xtest)


class Model:
    """Common base of DSMGP / PoE / gPoE / rBCM (DeepStructuredMixtures.jl:108-130): root + overlap D + gpmap."""
    predict_mode = nat.PREDICT_DSMGP

    def __init__(self, root: Node, x: np.ndarray, y: np.ndarray, kernels: List[KernelFunction], logNoise: float,
                 **handle_opts):
        self.root = root
        self.x = np.asarray(x, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64)
        self.kernels = kernels
        self.flat, self.leaves = flatten(root)
        self._D: Optional[np.ndarray] = None
        obs = [lf.obs for lf in self.leaves]
        yc = [self.y[lf.obs - 1] - lf.mean for lf in self.leaves]              # apply_subtract! means.jl:11-14
        self.handle = Handle(self.x, obs, yc, [lf.mean for lf in self.leaves], [lf.kernelid - 1 for lf in self.leaves],
                             kernels, self.flat, **handle_opts)
        self.setparams_(np.concatenate([np.concatenate([k.logl, [k.logs if k.type in (nat.ISO_SE, nat.ARD_SE) else 0.0,
                                                                  logNoise]]) for k in kernels]))

    # gpmap (BiDict, treeStructure.jl:421-426): leaf number <-> node id
    @property
    def gpmap(self):
        return {lf.id: i for i, lf in enumerate(self.leaves)}

    @property
    def D(self) -> np.ndarray:
        """Overlap matrix (getOverlap fit.jl:12-39, built at treeStructure.jl:428-431); computed lazily, on the device
        (`dsmgp_overlap`; `structure.getOverlap` is the host restatement of the reference's bit-set loop)."""
        if self._D is None:
            from .linalg import overlap_matrix
            self._D = overlap_matrix(self.x.shape[0], [lf.obs for lf in self.leaves],
                                     [lf.kernelid - 1 for lf in self.leaves], self.flat)
        return self._D

    @property
    def nparams(self) -> int:
        return self.handle.nparams

    def setparams_(self, hyp: Sequence[float]):
        """setparams!(spn, hyp) optimize.jl:188-198: one global theta for every leaf (per-kernel slices)."""
        hyp = np.asarray(hyp, dtype=np.float64)
        self.handle.set_params(hyp)
        self._mirror_params(hyp)

    def _mirror_params(self, hyp: np.ndarray):
        """Keep the host-side kernel objects in step with the device parameters (no device work)."""
        c = 0
        for k in self.kernels:
            nl = k.logl.size
            k.logl[:] = hyp[c:c + nl]
            if k.type in (nat.ISO_SE, nat.ARD_SE):
                k.logs = float(hyp[c + nl])
            c += nl + 2
        for lf in self.leaves:
            kk = self.kernels[lf.kernelid - 1]
            lf.kernel.logl[:] = kk.logl
            lf.kernel.logs = kk.logs
        c = 0
        noises = []
        for k in self.kernels:
            noises.append(float(hyp[c + k.logl.size + 1])); c += k.nparams
        for lf in self.leaves:
            lf.logNoise = noises[lf.kernelid - 1]

    def close(self):
        self.handle.close()


class DSMGP(Model):
    predict_mode = nat.PREDICT_DSMGP


class PoE(Model):
    predict_mode = nat.PREDICT_POE


class gPoE(Model):
    predict_mode = nat.PREDICT_GPOE


class rBCM(Model):
    predict_mode = nat.PREDICT_RBCM


# ---------------------------------------------------------------------------------------------
def build(x, y, K: int, V: int, eps: float, M: int, D: int, kernel, meanFun, logNoise: float, useSum: bool,
          cls=DSMGP, rng=None, fit: bool = True, **handle_opts) -> Model:
    """build(...) treeStructure.jl:405-437.  NOTE the reference's argument swap (:408-418): positional `K` (3rd)
    lands in DSMGPConfig.V (children per sum node) and `V` (4th) in DSMGPConfig.K (splits per split node)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    y = np.asarray(y, dtype=np.float64)
    rng = np.random.default_rng() if rng is None else (np.random.default_rng(rng) if isinstance(rng, int) else rng)
    kernels = [k.copy() for k in kernel] if isinstance(kernel, (list, tuple)) else kernel.copy()
    config = DSMGPConfig(meanFun, kernels, float(logNoise), int(M), int(V), int(K), int(D), float(eps), bool(useSum))
    root = buildTree(x, y, config, rng)
    klist = kernels if isinstance(kernels, list) else [kernels]
    model = cls(root, x, y, klist, float(logNoise), **handle_opts)
    if fit:
        fit_(model)                                            # treeStructure.jl:434
    return model


def buildDSMGP(x, y, V: int, K: int, *, eps: float = 0.5, M: int = 30, D: int = 2, kernel=None, meanFun=None,
               logNoise: float = 1.0, sum: bool = True, rng=None, **kw) -> DSMGP:
    """buildDSMGP(x, y, V, K; ϵ, M, D, kernel, meanFun, logNoise, sum) treeStructure.jl:328-339.
    V = children per sum node, K = splits per split node (README.md:47-51)."""
    kernel = IsoSE(1.0, 1.0) if kernel is None else kernel
    return build(x, y, V, K, eps, M, D, kernel, meanFun, logNoise, sum, cls=DSMGP, rng=rng, **kw)


def buildPoE(x, y, V: int, *, eps: float = 0.0, M: int = 30, D: int = 2, kernel=None, meanFun=None,
             logNoise: float = 1.0, generalized: bool = False, rng=None, **kw) -> Model:
    """buildPoE(x, y, V; ...) treeStructure.jl:360-371 (V = splits per split node)."""
    kernel = IsoSE(1.0, 1.0) if kernel is None else kernel
    return build(x, y, 1, V, eps, M, D, kernel, meanFun, logNoise, False, cls=gPoE if generalized else PoE, rng=rng, **kw)


def buildBCM(x, y, V: int, *, eps: float = 0.0, M: int = 30, D: int = 2, kernel=None, meanFun=None,
             logNoise: float = 1.0, robust: bool = False, rng=None, **kw) -> rBCM:
    """buildBCM(x, y, V; ...) treeStructure.jl:392-403 (always returns an rBCM, like the reference)."""
    kernel = IsoSE(1.0, 1.0) if kernel is None else kernel
    return build(x, y, 1, V, eps, M, D, kernel, meanFun, logNoise, False, cls=rBCM, rng=rng, **kw)


def GaussianProcess(x, y, *, mean: Optional[ConstMean] = None, kernel: Optional[KernelFunction] = None,
                    logNoise: float = math.log(7.0), run_cholesky: bool = False, **kw) -> LeafGP:
    """GaussianProcess(x, y; mean, kernel, logNoise, run_cholesky) gaussianprocess.jl:50-80: one exact GP =
    a model whose region graph is a single leaf."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    y = np.asarray(y, dtype=np.float64)
    kernel = IsoSE(0.0, 0.0) if kernel is None else kernel.copy()
    m = float(np.mean(y)) if mean is None else float(mean.m)
    leaf = GPNode(obs=np.arange(1, len(y) + 1, dtype=np.int64), lb=np.full(x.shape[1], -np.inf),
                  ub=np.full(x.shape[1], np.inf), nobs=len(y), kernelid=1, mean=m, kernel=kernel, logNoise=logNoise)
    model = DSMGP(leaf, x, y, [kernel.copy()], float(logNoise), **kw)
    gp = LeafGP(model, 0)
    if run_cholesky:
        update_cholesky_(gp)
    return gp


# ---------------------------------------------------------------------------------------------
def _model_of(obj) -> Model:
    return obj.model if isinstance(obj, LeafGP) else obj


def fit_(model: Model, tau: float = 0.05) -> float:
    """fit!(model; τ) fit.jl:67-122 with the model's overlap matrix D: the shared Cholesky (identical experts factored
    once, common leading block rows reused; every result is the exact factor of update_cholesky!).  The plan stays in the
    handle for the following evaluations, like train! calling fit!(spn, D, gpmap) every iteration.  Returns device seconds."""
    if model.handle.world > 1 or len(model.leaves) == 1 or len(model.leaves) > 8192:
        _, sec = model.handle.fit()          # (a dense 8192 x 8192 overlap matrix is where the plan stops paying for itself)
    else:
        _, sec = model.handle.fit(model.D, tau)
    return sec


def fit_naive_(model: Model) -> float:
    """fit_naive!(spn) fit.jl:294-304: every expert factored on its own."""
    _, sec = model.handle.fit()
    return sec


def update_cholesky_(gp: LeafGP) -> LeafGP:                   # gaussianprocess.jl:82-108
    gp.model.handle.fit()
    return gp


def mll(obj: Union[Model, LeafGP]) -> float:
    """mll(model) optimize.jl:18-25 / mll(gp) gaussianprocess.jl:163 (requires a preceding fit_)."""
    m = _model_of(obj)
    return float(m.handle.lml()[m.flat.root])


def mll_nodes(model: Model) -> np.ndarray:
    """mll!(spn, L) optimize.jl:27-39: the per-node table, indexed by node id."""
    return model.handle.lml()


def grad_mll(obj: Union[Model, LeafGP], leaf_scale=None) -> np.ndarray:
    """updategradients! + ∇mll! (fit.jl:306-311, optimize.jl:42-150; single GP: gaussianprocess.jl:185-217)."""
    return _model_of(obj).handle.grad(leaf_scale)


def evaluate(model: Model, hyp=None, leaf_scale=None) -> Tuple[float, np.ndarray]:
    """One LML+gradient evaluation (optimisers.jl:43-77 minus the Flux step) in a single library call."""
    lml, g = model.handle.eval(hyp, leaf_scale)
    if hyp is not None:
        model._mirror_params(np.asarray(hyp, dtype=np.float64))
    return lml, g


def update_(model: Model) -> float:
    """update!(model) common.jl:323-334: writes the posterior log-weights into every sum node, returns z."""
    lw, z = model.handle.update_weights()
    ft = model.flat

    def rec(n: Node):
        if isinstance(n, GPNode):
            return
        if isinstance(n, GPSumNode):
            n.logweights = [float(v) for v in lw[ft.child_ptr[n.id]:ft.child_ptr[n.id + 1]]]
        for c in n.children:
            rec(c)

    rec(model.root)
    return z


def _write_logweights(model: Model, lw: np.ndarray):
    ft = model.flat

    def rec(n: Node):
        if isinstance(n, GPNode):
            return
        if isinstance(n, GPSumNode):
            n.logweights = [float(v) for v in lw[ft.child_ptr[n.id]:ft.child_ptr[n.id + 1]]]
        for c in n.children:
            rec(c)

    rec(model.root)


def infer_(model: Model) -> float:
    """infer!(model) common.jl:336-355: kernel-mixture sum nodes keep posterior weights, the other sum nodes are reset to
    uniform after their evidence is computed; returns z."""
    lw, z = model.handle.infer()
    _write_logweights(model, lw)
    return z


def reset_weights_(model: Model) -> None:
    """reset_weights!(model) common.jl:357-363."""
    _write_logweights(model, model.handle.reset_weights())


def predict(model: Model, x) -> Tuple[np.ndarray, np.ndarray]:
    """predict(model, x) common.jl:304-307 -> (mean, variance)."""
    return model.handle.predict(x, model.predict_mode)


def prediction(gp: LeafGP, xtest):
    return gp.prediction(xtest)


def leftGP(node_or_model):
    """leftGP common.jl:124-127: the left-most expert (a list of experts for a kernel mixture)."""
    model = node_or_model
    n = model.root
    while not isinstance(n, GPNode):
        if isinstance(n, GPSumNode) and n.kernel_mixture:
            return [LeafGP(model, c.leaf_index) for c in n.children]
        n = n.children[0]
    return LeafGP(model, n.leaf_index)


def rightGP(node_or_model):
    model = node_or_model
    n = model.root
    while not isinstance(n, GPNode):
        if isinstance(n, GPSumNode) and n.kernel_mixture:
            return [LeafGP(model, c.leaf_index) for c in n.children]
        n = n.children[-1]
    return LeafGP(model, n.leaf_index)


def params(gp: Union[LeafGP, List[LeafGP]], logscale: bool = False):
    if isinstance(gp, list):
        return [g.params(logscale) for g in gp]
    return gp.params(logscale)


def setparams_(obj: Union[Model, LeafGP], hyp):
    if isinstance(obj, LeafGP):
        obj.setparams_(hyp)
    else:
        obj.setparams_(hyp)


def stats(model: Model) -> dict:
    """stats(node) common.jl:365-395 (structural counts only)."""
    out = {"gps": 0, "ndata": [], "sumnodes": 0, "slitnodes": 0}

    def rec(n: Node):
        if isinstance(n, GPNode):
            out["gps"] += 1; out["ndata"].append(n.nobs); return
        if isinstance(n, GPSumNode):
            out["sumnodes"] += 1
        else:
            out["slitnodes"] += 1
        for c in n.children:
            rec(c)

    rec(model.root)
    return out
