"""Training drivers (host orchestration): train_ / finetune_ and Flux-style optimisers.

Mirrors /root/reference/src/optimisers.jl:4-145 and src/finetuning.jl:3-88.  These are the CALLERS of the hot
path: every iteration is one `dsmgp_eval` (setparams! -> fit! -> mll! -> updategradients! -> ∇mll!).

Flux semantics that matter: `Flux.Optimise.apply!(opt, x, Δ)` turns Δ into the step in place
with optimiser state keyed by the IDENTITY of `x`; the reference then rebinds `hyp += grad`, so the state is
fresh every iteration and an ADAM step degenerates to η·Δ/(|Δ|+ϵ).  `state_by_identity=True` (default)
reproduces that; `False` keeps one state per optimiser (textbook behaviour).  The update is gradient ASCENT.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from .model import LeafGP, Model, fit_, leftGP, update_cholesky_
from .structure import GPSumNode


class _Optimiser:
    def __init__(self, state_by_identity: bool = True):
        self.state_by_identity = state_by_identity
        self._state: Dict[int, tuple] = {}

    def _get(self, x: np.ndarray, init):
        key = id(x) if self.state_by_identity else 0
        ent = self._state.get(key)
        if ent is None or (self.state_by_identity and ent[0] is not x):
            ent = (x, init())
            self._state[key] = ent
        return ent[1]

    def _put(self, x: np.ndarray, st):
        self._state[id(x) if self.state_by_identity else 0] = (x, st)


class Descent(_Optimiser):
    def __init__(self, eta: float = 0.1, **kw):
        super().__init__(**kw); self.eta = eta

    def apply_(self, x, delta):
        delta *= self.eta
        return delta


class ADAM(_Optimiser):
    """Flux.ADAM(η=0.001, β=(0.9, 0.999)), ϵ = 1e-8."""

    def __init__(self, eta: float = 0.001, beta=(0.9, 0.999), **kw):
        super().__init__(**kw); self.eta, self.beta = eta, beta

    def apply_(self, x, delta):
        b1, b2 = self.beta
        mt, vt, bp = self._get(x, lambda: (np.zeros_like(x), np.zeros_like(x), [b1, b2]))
        mt = b1 * mt + (1 - b1) * delta
        vt = b2 * vt + (1 - b2) * delta ** 2
        delta[:] = mt / (1 - bp[0]) / (np.sqrt(vt / (1 - bp[1])) + 1e-8) * self.eta
        self._put(x, (mt, vt, [bp[0] * b1, bp[1] * b2]))
        return delta


class RMSProp(_Optimiser):
    """Flux.RMSProp(η=0.001, ρ=0.9), ϵ = 1e-8."""

    def __init__(self, eta: float = 0.001, rho: float = 0.9, **kw):
        super().__init__(**kw); self.eta, self.rho = eta, rho

    def apply_(self, x, delta):
        acc = self._get(x, lambda: np.zeros_like(x))
        acc = self.rho * acc + (1 - self.rho) * delta ** 2
        delta[:] = delta * (self.eta / (np.sqrt(acc) + 1e-8))
        self._put(x, acc)
        return delta


def _current_hyp(model: Model) -> np.ndarray:
    gp = leftGP(model)
    gps = gp if isinstance(gp, list) else [gp]
    out = []
    for g in gps:
        l, s, n = g.params(logscale=True)
        out.append(np.concatenate([np.atleast_1d(l), [s, n]]))
    return np.concatenate(out)


def train_(model, optim=None, *, iterations: int = 10_000, lam: float = 0.05, randinit: bool = True,
           earlystop: int = 10, rng=None, callback=None):
    """train!(model, optim; iterations, λ, randinit, earlystop) optimisers.jl:4-87.
    Returns (model, ℓ) with ℓ the LML trace.  For a single `GaussianProcess` see `train_gp_`."""
    if isinstance(model, LeafGP):
        return train_gp_(model, optim=optim, iterations=iterations, lam=lam, rng=rng)
    optim = ADAM() if optim is None else optim
    rng = np.random.default_rng() if rng is None else (np.random.default_rng(rng) if isinstance(rng, int) else rng)
    n = model.nparams
    hyp = rng.standard_normal(n) if randinit else _current_hyp(model)      # :18
    if callback is None and type(optim) in (Descent, ADAM, RMSProp):
        # the loop itself runs inside the library (dsmgp_train): no per-iteration round trip
        kind = {Descent: 0, ADAM: 1, RMSProp: 2}[type(optim)]
        b1, b2 = (optim.beta if kind == 1 else ((optim.rho, 0.0) if kind == 2 else (0.0, 0.0)))
        hyp, ell = model.handle.train(hyp, kind, optim.eta, b1, b2, optim.state_by_identity, iterations, lam, earlystop)
        model._mirror_params(hyp)
        return model, ell
    ell = np.zeros(iterations)
    c = 0
    delta = np.inf
    for it in range(iterations):
        lml, grad = model.handle.eval(hyp)                                  # :43-49, 68-77
        ell[it] = lml
        delta = abs(ell[it] - np.mean(ell[it - 9:it])) if it >= 10 else np.inf   # :53
        c = c + 1 if delta < lam else 0                                     # :57-61
        if callback is not None:
            callback(it, lml, grad, hyp)
        if c >= earlystop:                                                  # :63-66 (returns before the update)
            model._mirror_params(hyp)
            return model, ell[:it + 1]
        optim.apply_(hyp, grad)                                             # :78
        hyp = hyp + grad                                                    # :79 (rebinding; gradient ASCENT)
    model.setparams_(hyp)                                                   # :82-83
    fit_(model)
    return model, ell


def train_gp_(gp: LeafGP, *, optim=None, iterations: int = 10_000, lam: float = 0.1, rng=None):
    """train!(gp; iterations, optim, λ) optimisers.jl:89-145 (single exact GP, rolls back on NaN)."""
    optim = RMSProp() if optim is None else optim
    rng = np.random.default_rng() if rng is None else (np.random.default_rng(rng) if isinstance(rng, int) else rng)
    model = gp.model
    n = model.nparams
    hyp = rng.standard_normal(n)
    oldhyp = hyp
    ell = np.zeros(iterations)
    for it in range(iterations):
        lml, grad = model.handle.eval(hyp)
        ell[it] = lml
        if np.isnan(lml):                                                   # :115-119
            gp.setparams_(oldhyp); update_cholesky_(gp)
            return gp, ell[:it + 1]
        delta = abs(ell[it] - np.mean(ell[it - 9:it])) if it >= 10 else np.inf
        if delta < lam:                                                     # :125-128
            model._mirror_params(hyp)
            return gp, ell[:it + 1]
        oldhyp = hyp.copy()
        optim.apply_(hyp, grad)
        hyp = hyp + grad
    gp.setparams_(hyp); update_cholesky_(gp)
    return gp, ell


def finetune_(model: Model, optim=None, *, iterations: int = 1000, lam: float = 0.5, batch: int = 256,
              strict_reference: bool = False):
    """finetune!(model, optim; iterations, λ) finetuning.jl:3-88: per-leaf hyper-parameters.  For every leaf g the
    WHOLE model is evaluated under θ_g and the leaf gradients are weighted by the overlap row D[g,:]
    (optimize.jl:92-102; the diagonal of D is 0, so a leaf's own gradient has weight 0).

    Kernel mixtures (`KernelFunction[...]`): the reference indexes a single kernel's 3-vector as if it held every kernel's
    parameters and throws a BoundsError (finetuning.jl:41; `strict_reference=True` reproduces that).  The
    library defines the semantics the code was reaching for: every expert keeps ITS OWN kernel's theta; the evaluation anchored at
    expert g sets, for every kernel k, theta_k of the whole model to the theta of the kernel-k expert of g's REGION (the siblings
    under g's kernel-mixture sum node, treeStructure.jl:258-286), and g is updated from the slice of the gradient that belongs to
    its own kernel (optimize.jl:76-89).  BASELINE.json config 4 ("gPoE warm-start then DSMGP finetune!") runs this way.

    The L evaluations of one iteration are independent (hyp[g] is only updated from its own gradient), so they go to the
    device as ONE `dsmgp_finetune_eval` call; `batch` bounds how many anchors share a call."""
    optim = ADAM() if optim is None else optim
    nk = len(model.kernels)
    if nk != 1 and strict_reference:
        raise IndexError("finetune! with a kernel vector is a BoundsError in the reference (finetuning.jl:41)")
    L = len(model.leaves)
    D = model.D
    hyp = [model.handle.get_leaf_params(g).copy() for g in range(L)]         # :24
    # region of every expert: the experts (one per kernel, in kernel order) under its kernel-mixture sum node, or itself
    koff = np.concatenate([[0], np.cumsum([k.nparams for k in model.kernels])])
    region = {}
    if nk > 1:
        def rec(n):
            if isinstance(n, GPSumNode) and n.kernel_mixture:
                ids = [c.leaf_index for c in n.children]
                for g in ids:
                    region[g] = ids
                return
            for c in getattr(n, "children", []):
                rec(c)
        rec(model.root)

    def full_theta(g):
        return hyp[g] if nk == 1 else np.concatenate([hyp[s] for s in region[g]])

    def own_slice(g, grad):
        if nk == 1:
            return grad
        k = model.leaves[g].kernelid - 1
        return grad[koff[k]:koff[k + 1]]
    ell = np.zeros(iterations)
    c = 0
    for it in range(iterations):
        l = 0.0
        for g0 in range(0, L, batch):
            gs = list(range(g0, min(L, g0 + batch)))
            leaf_lml, grads, _ = model.handle.finetune_eval(gs, np.stack([full_theta(g) for g in gs]), D)   # :41-54
            for i, g in enumerate(gs):
                l += leaf_lml[i]                                             # :51
                grad = own_slice(g, grads[i]).copy()
                optim.apply_(hyp[g], grad)
                hyp[g] = hyp[g] + grad
        ell[it] = l
        delta = abs(ell[it] - np.mean(ell[it - 9:it])) if it >= 10 else np.inf
        c = c + 1 if delta < lam else 0
        if c >= 10:
            break
    for g in range(L):                                                       # :75-84
        model.handle.set_leaf_params(g, hyp[g])
    fit_(model)
    return model, ell
