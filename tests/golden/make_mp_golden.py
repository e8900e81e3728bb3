#!/usr/bin/env python
"""INDEPENDENT pin of the hot path: 50-digit mpmath evaluation of tiny models straight from the Julia formulas.

This script does NOT import oracle/dsm_oracle.py (nor NumPy's linear algebra, SciPy or the product): every quantity is
computed with mpmath at 50 digits from the reference's source lines cited below, with its own Cholesky, triangular
solves, inverse, tree passes and mixing.  The reference ships no tests or golden vectors and Julia is not installed, so
this is the pin that is not the oracle itself: tests/test_oracle.py checks the oracle against these vectors and
tests/test_gpu_parity.py checks libdsmgp (through the C ABI) against them.

    python tests/golden/make_mp_golden.py        # rewrites tests/golden/mp_golden.json

Models (N = 24 points, D = 2; every leaf has 12 points):
  dsmgp_<kernel>   root sum node (2 children) -> split node on dim 0 / dim 1 (2 intervals each) -> 4 leaf GPs,
                   kernel in {IsoSE, ArdSE, IsoLinear, ArdLinear}
  dsmgp_mixture    the same regions, every region a kernel-mixture sum node over [IsoSE, IsoLinear] GPs (treeStructure.jl:258-286)
  poe_isose        root split node on dim 0 with 2 leaf GPs: predictPoE / predictgPoE / predictrBCM
Quantities: per-leaf mll (gaussianprocess.jl:163), per-leaf gradients as written (kernels.jl:85-99,146-164,196-200) and
mathematical (true d mll / d theta), the per-node table of mll! (optimize.jl:27-39), the model gradient of the down-pass
(optimize.jl:42-89) for both leaf-gradient modes and with a finetune weight row (:92-102), update! / infer! weights and z
(common.jl:323-355), predictions (common.jl:134-313; gaussianprocess.jl:110-137).
"""
import json
import os
import random

import mpmath as mp

mp.mp.dps = 50
HERE = os.path.dirname(os.path.abspath(__file__))
EPS = mp.mpf("1e-8")                      # DeepStructuredMixtures.jl:27
LOG2PI = mp.log(2 * mp.pi)                # StatsFuns.log2π

ISO_SE, ARD_SE, ISO_LINEAR, ARD_LINEAR = 0, 1, 2, 3
LEAF, SPLIT, SUM, KSUM = 0, 1, 2, 3


# ---- kernels.jl -----------------------------------------------------------------------------------------------------
def nl_of(ktype, D):
    return 1 if ktype in (ISO_SE, ISO_LINEAR) else D


def kval(ktype, th, a, b):
    """kernelmatrix entry K(a, b); th = [logl..., logsigma, logNoise] (floats, used exactly)."""
    D = len(a)
    nl = nl_of(ktype, D)
    ell = [mp.e ** mp.mpf(th[d]) for d in range(nl)]
    if ktype == ISO_SE:        # kernels.jl:21-26,78,83: v * exp(-0.5 * (|a-b|^2 / l^2))
        v = mp.e ** (2 * mp.mpf(th[nl]))
        r2 = sum((mp.mpf(a[d]) - mp.mpf(b[d])) ** 2 for d in range(D))
        return v * mp.e ** (-mp.mpf("0.5") * (r2 / ell[0] ** 2))
    if ktype == ARD_SE:        # kernels.jl:31-49: ADDITIVE over dimensions
        v = mp.e ** (2 * mp.mpf(th[nl]))
        return v * sum(mp.e ** (-mp.mpf("0.5") * ((mp.mpf(a[d]) - mp.mpf(b[d])) ** 2 / ell[d] ** 2)) for d in range(D))
    if ktype == ISO_LINEAR:    # kernels.jl:189,194: (a . b) / l^2, variance fixed to 1 (:181)
        return sum(mp.mpf(a[d]) * mp.mpf(b[d]) for d in range(D)) / ell[0] ** 2
    return sum(mp.mpf(a[d]) * mp.mpf(b[d]) / ell[d] ** 2 for d in range(D))     # ArdLinear: SURVEY App. A.2


def chol_lower(A):
    n = A.rows
    L = mp.zeros(n)
    for j in range(n):
        s = A[j, j] - sum(L[j, k] ** 2 for k in range(j))
        L[j, j] = mp.sqrt(s)
        for i in range(j + 1, n):
            L[i, j] = (A[i, j] - sum(L[i, k] * L[j, k] for k in range(j))) / L[j, j]
    return L


def fwd(L, b):          # L \ b
    n = L.rows
    z = [mp.mpf(0)] * n
    for i in range(n):
        z[i] = (b[i] - sum(L[i, k] * z[k] for k in range(i))) / L[i, i]
    return z


def bwd(L, z):          # L' \ z
    n = L.rows
    a = [mp.mpf(0)] * n
    for i in reversed(range(n)):
        a[i] = (z[i] - sum(L[k, i] * a[k] for k in range(i + 1, n))) / L[i, i]
    return a


class GP:
    """gaussianprocess.jl:14-226 for one leaf."""

    def __init__(self, ktype, X, y, mean):
        self.ktype, self.X, self.mean = ktype, X, mp.mpf(mean)
        self.y = [mp.mpf(v) - self.mean for v in y]              # apply_subtract! means.jl:11-14
        self.n, self.D = len(X), len(X[0])

    def fit(self, th):
        n = self.n
        self.th = th
        nl = nl_of(self.ktype, self.D)
        self.eta = mp.e ** (2 * mp.mpf(th[nl + 1]))                   # getnoise :39
        self.K = mp.matrix(n, n)
        for i in range(n):
            for j in range(n):
                self.K[i, j] = kval(self.ktype, th, self.X[i], self.X[j])
        F = self.K.copy()
        for i in range(n):
            F[i, i] += self.eta + EPS                                  # :93-98
        self.L = chol_lower(F)                                         # :101
        self.alpha = bwd(self.L, fwd(self.L, self.y))                  # :105
        logdet = 2 * sum(mp.log(self.L[i, i]) for i in range(n))
        self.lml = -(sum(self.y[i] * self.alpha[i] for i in range(n)) + logdet + LOG2PI * n) / 2     # :163
        Finv = mp.zeros(n)
        for c in range(n):
            e = [mp.mpf(1) if i == c else mp.mpf(0) for i in range(n)]
            col = bwd(self.L, fwd(self.L, e))
            for i in range(n):
                Finv[i, c] = col[i]
        self.W = mp.matrix(n, n)                                       # ααinvcK! :219-226
        for i in range(n):
            for j in range(n):
                self.W[i, j] = self.alpha[i] * self.alpha[j] - Finv[i, j]
        return self

    def grads(self, as_written):
        """[dl..., dsigma, dnoise] (gaussianprocess.jl:206-217)."""
        n, D, th, kt = self.n, self.D, self.th, self.ktype
        nl = nl_of(kt, D)
        W, K = self.W, self.K
        trW = sum(W[i, i] for i in range(n))
        trWK = sum(W[i, j] * K[j, i] for i in range(n) for j in range(n))
        dnoise = self.eta * trW                                        # :176
        ell = [mp.e ** mp.mpf(th[d]) for d in range(nl)]
        s = mp.e ** mp.mpf(th[nl]) if kt in (ISO_SE, ARD_SE) else mp.mpf(1)
        v = s * s
        dl = [mp.mpf(0)] * nl
        if kt == ISO_SE:
            # kernels.jl:85-99: K <- sigma K ; dsigma = 0.5 tr(precomp * 2K) ; K .*= P / l^2 ; dl = 0.5 tr(precomp * K)
            acc = mp.mpf(0)
            for i in range(n):
                for j in range(n):
                    r2 = sum((mp.mpf(self.X[i][d]) - mp.mpf(self.X[j][d])) ** 2 for d in range(D))
                    acc += W[i, j] * K[j, i] * r2 / ell[0] ** 2
            fac = s if as_written else mp.mpf(1)
            dsig = fac * trWK
            dl[0] = mp.mpf("0.5") * fac * acc
        elif kt == ARD_SE:
            fac = s if as_written else mp.mpf(1)
            dsig = fac * trWK                                          # kernels.jl:154-157
            for d in range(D):
                if as_written:
                    dl[d] = mp.mpf(0)                                  # :161  tr((precomp*K) .* (p/ls[d])), diag(p) == 0
                else:
                    acc = mp.mpf(0)
                    for i in range(n):
                        for j in range(n):
                            p = (mp.mpf(self.X[i][d]) - mp.mpf(self.X[j][d])) ** 2
                            acc += W[i, j] * v * mp.e ** (-mp.mpf("0.5") * (p / ell[d] ** 2)) * p / ell[d] ** 2
                    dl[d] = mp.mpf("0.5") * acc
        elif kt == ISO_LINEAR:
            dsig = mp.mpf(0)                                           # getgradients :201
            dl[0] = -trWK                                              # :198  0.5 tr(precomp * -2K)
        else:
            dsig = mp.mpf(0)
            for d in range(D):                                         # SURVEY App. A.4 (reference method is broken, Q5)
                dl[d] = -sum(W[i, j] * mp.mpf(self.X[j][d]) * mp.mpf(self.X[i][d]) / ell[d] ** 2 for i in range(n) for j in range(n))
        return dl + [dsig, dnoise]

    def predict(self, xt):
        """prediction(gp, xtest) gaussianprocess.jl:110-137 -> (mu, diag Sigma)."""
        n = self.n
        knt = [kval(self.ktype, self.th, self.X[i], xt) for i in range(n)]
        mu = self.mean + sum(knt[i] * self.alpha[i] for i in range(n))
        V = fwd(self.L, knt)
        var = kval(self.ktype, self.th, xt, xt) - sum(vv * vv for vv in V) + self.eta      # no 1e-8 here (:123-126)
        return mu, var


# ---- tree -------------------------------------------------------------------------------------------------------------
class Node:
    def __init__(self, typ, children=(), split=None, gp=None, kid=0, obs=None):
        self.type, self.children, self.split, self.gp, self.kid, self.obs = typ, list(children), split, gp, kid, obs
        self.id = -1
        self.leaf = -1
        self.logw = None


def number(root):
    """children before parents; leaves numbered in getLeaves order (fit.jl:9-10)."""
    order, leaves = [], []

    def rec(n):
        for c in n.children:
            rec(c)
        n.id = len(order)
        order.append(n)
        if n.type == LEAF:
            n.leaf = len(leaves)
            leaves.append(n)

    def leaves_dfs(n):
        if n.type == LEAF:
            return [n]
        out = []
        for c in n.children:
            out += leaves_dfs(c)
        return out
    rec(root)
    for i, lf in enumerate(leaves_dfs(root)):
        lf.leaf = i
    return order, leaves_dfs(root)


def lse(vals):                      # StatsFuns.logsumexp / common.jl:309-313
    m = max(vals)
    return m + mp.log(sum(mp.e ** (v - m) for v in vals))


def mll_up(n, ell):                 # optimize.jl:27-39
    if n.type == LEAF:
        v = n.gp.lml
    elif n.type == SPLIT:
        v = sum(mll_up(c, ell) for c in n.children)
    else:
        K = len(n.children)
        v = lse([-mp.log(K) + mll_up(c, ell) for c in n.children])
    ell[n.id] = v
    return v


def grad_down(n, dpar, lrho, ell, logS, grad, off, leafg, Drow):     # optimize.jl:42-150
    if n.type == LEAF:
        w = mp.e ** (-logS + lrho + ell[n.id] + dpar)
        if Drow is not None:
            w = w * mp.mpf(Drow[n.leaf])
        g = leafg[n.leaf]
        for k in range(len(g)):
            grad[off + k] += g[k] * w
    elif n.type == SPLIT:
        for c in n.children:
            grad_down(c, dpar + (ell[n.id] - ell[c.id]), lrho, ell, logS, grad, off, leafg, Drow)
    elif n.type == SUM:
        K = len(n.children)
        for c in n.children:
            grad_down(c, -mp.log(K) + dpar, mp.log(K) + lrho, ell, logS, grad, off, leafg, Drow)
    else:                                                              # kernel mixture :76-89
        c0 = 0
        for c in n.children:
            grad_down(c, dpar, lrho, ell, logS, grad, off + c0, leafg, Drow)
            c0 += len(leafg[c.leaf])


def update_w(n, infer=False):       # common.jl:323-334 / 336-355
    if n.type == LEAF:
        return n.gp.lml
    if n.type == SPLIT:
        return sum(update_w(c, infer) for c in n.children)
    K = len(n.children)
    lw = [-mp.log(K) + update_w(c, infer) for c in n.children]
    z = lse(lw)
    if infer and n.type == SUM:
        n.logw = [-mp.log(K)] * K                                      # :347-353
    else:
        n.logw = [v - z for v in lw]
    return z


def getchild(n, xt):                # common.jl:101-122
    d = n.split[0][0]
    for k, (_, s) in enumerate(n.split):
        ok = (xt[d] <= s) if k == 0 else ((xt[d] <= s) and (xt[d] > n.split[k - 1][1]))
        if ok:
            return k
    raise RuntimeError("point outside every interval")


def minpredict(n, xt):              # common.jl:151-173
    if n.type == LEAF:
        return n.gp.predict(xt)[0]
    if n.type == SPLIT:
        return minpredict(n.children[getchild(n, xt)], xt)
    return min(minpredict(c, xt) for c in n.children)


def predict_rec(n, xt, mumin):      # common.jl:134-143,181-196,275-292
    if n.type == LEAF:
        mu, var = n.gp.predict(xt)
        if var <= 0:
            var = EPS
        return mp.log(mu - mumin), mp.log(mu ** 2), mp.log(var)
    if n.type == SPLIT:
        return predict_rec(n.children[getchild(n, xt)], xt, mumin)
    parts = [predict_rec(c, xt, mumin) for c in n.children]
    return tuple(lse([parts[k][q] + n.logw[k] for k in range(len(parts))]) for q in range(3))


def predict_dsmgp(root, xt):        # common.jl:294-302
    mumin = minpredict(root, xt) - 1
    lm, lm2, ls = predict_rec(root, xt, mumin)
    mu = mp.e ** lm + mumin
    return mu, mp.e ** ls + (mp.e ** lm2 - mu ** 2)


def poe_rec(n, xt):                 # common.jl:145-149,198-208
    if n.type == LEAF:
        mu, var = n.gp.predict(xt)
        return mu, 1 / var
    t = mp.mpf(0); m = mp.mpf(0)
    for c in n.children:
        m_, t_ = poe_rec(c, xt)
        t += t_; m += t_ * m_
    return m / t, t


def leftgp(n):
    while n.type != LEAF:
        n = n.children[0]
    return n.gp


# ---- data and models ------------------------------------------------------------------------------------------------
def f17(v):
    return float(mp.nstr(v, 25))


def make_data():
    rng = random.Random(20240607)
    N, D = 24, 2
    X = [[round(rng.random(), 6) for _ in range(D)] for _ in range(N)]
    y = []
    for i in range(N):
        y.append(round(float(mp.sin(2 * mp.pi * (mp.mpf("0.7") * X[i][0] - mp.mpf("0.4") * X[i][1]))) + 0.1 * rng.gauss(0, 1), 6))
    T = [[round(0.05 + 0.9 * rng.random(), 6) for _ in range(D)] for _ in range(6)]
    return X, y, T


def region_split(X, d):
    vals = sorted(X[i][d] for i in range(len(X)))
    s = 0.5 * (vals[len(vals) // 2 - 1] + vals[len(vals) // 2])
    lo = [i for i in range(len(X)) if X[i][d] <= s]
    hi = [i for i in range(len(X)) if X[i][d] > s]
    return s, lo, hi


def build(X, y, kernels, poe=False):
    """kernels: list of kernel types (one: plain leaves; two: kernel-mixture sum per region)."""
    def region(obs):
        def leaf(kid):
            yy = [y[i] for i in obs]
            m = float(sum(mp.mpf(v) for v in yy) / len(yy))              # ConstMean(mean(y_leaf)) treeStructure.jl:271,292
            return Node(LEAF, gp=GP(kernels[kid], [X[i] for i in obs], yy, m), kid=kid, obs=[i + 1 for i in obs])
        if len(kernels) == 1:
            return leaf(0)
        return Node(KSUM, [leaf(k) for k in range(len(kernels))])
    dims = [0] if poe else [0, 1]
    splits = []
    for d in dims:
        s, lo, hi = region_split(X, d)
        splits.append(Node(SPLIT, [region(lo), region(hi)], split=[(d, s), (d, 2.0)]))      # last threshold = upperBound[d]
    root = splits[0] if poe else Node(SUM, splits)
    return root


def flat_of(root):
    order, leaves = number(root)
    node_type = [n.type for n in order]
    child_ptr, child_idx, split_ptr, split_val, split_dim, leaf_of_node = [0], [], [0], [], [], []
    for n in order:
        child_idx += [c.id for c in n.children]
        child_ptr.append(len(child_idx))
        if n.type == SPLIT:
            split_val += [s for _, s in n.split]
            split_dim.append(n.split[0][0])
        else:
            split_dim.append(-1)
        split_ptr.append(len(split_val))
        leaf_of_node.append(n.leaf if n.type == LEAF else -1)
    leaf_ptr, leaf_obs = [0], []
    for lf in leaves:
        leaf_obs += lf.obs
        leaf_ptr.append(len(leaf_obs))
    return dict(node_type=node_type, child_ptr=child_ptr, child_idx=child_idx, leaf_of_node=leaf_of_node, split_dim=split_dim,
                split_ptr=split_ptr, split_val=split_val, root=root.id, leaf_ptr=leaf_ptr, leaf_obs=leaf_obs,
                leaf_kernel_id=[lf.kid for lf in leaves], leaf_mean=[float(lf.gp.mean) for lf in leaves]), order, leaves


def theta_for(kt, D, which):
    base = {ISO_SE: [-0.3, 0.2, -0.9], ARD_SE: [-0.4, 0.1, 0.3, -0.8], ISO_LINEAR: [0.25, 0.0, -0.6], ARD_LINEAR: [0.2, -0.15, 0.0, -0.7]}[kt]
    alt = {ISO_SE: [0.1, -0.15, -1.2], ARD_SE: [0.05, -0.35, -0.2, -1.1], ISO_LINEAR: [-0.1, 0.0, -1.0], ARD_LINEAR: [-0.2, 0.3, 0.0, -1.05]}[kt]
    return base if which == 0 else alt


def run_model(name, X, y, T, kernels, poe=False):
    D = len(X[0])
    root = build(X, y, kernels, poe)
    flat, order, leaves = flat_of(root)
    out = dict(name=name, kernels=kernels, flat=flat, evals=[])
    Drow = [1.0, 0.0, 0.5, 0.25, 0.75, 0.0, 0.3, 0.6][:len(leaves)]
    for which in (0, 1):
        theta = []
        for kt in kernels:
            theta += theta_for(kt, D, which)
        off = [0]
        for kt in kernels:
            off.append(off[-1] + nl_of(kt, D) + 2)
        for lf in leaves:
            lf.gp.fit(theta[off[lf.kid]:off[lf.kid + 1]])
        ell = {}
        mll_up(root, ell)
        rec = dict(theta=theta, leaf_lml=[f17(lf.gp.lml) for lf in leaves], node_lml=[f17(ell[n.id]) for n in order],
                   leaf_alpha=[[f17(a) for a in lf.gp.alpha] for lf in leaves])
        for mode, aw in (("as_written", True), ("mathematical", False)):
            leafg = {lf.leaf: lf.gp.grads(aw) for lf in leaves}
            rec["leaf_grad_" + mode] = [[f17(v) for v in leafg[lf.leaf]] for lf in leaves]
            for tag, dr in (("", None), ("_finetune", Drow)):
                grad = [mp.mpf(0)] * off[-1]
                grad_down(root, mp.mpf(0), mp.mpf(0), ell, ell[root.id], grad, 0, leafg, dr)
                rec["grad_" + mode + tag] = [f17(v) for v in grad]
        rec["finetune_row"] = Drow
        if not poe:
            z = update_w(root, infer=False)
            rec["update_z"] = f17(z)
            rec["update_logw"] = [[f17(v) for v in n.logw] if n.type >= SUM else [] for n in order]
            preds = [predict_dsmgp(root, xt) for xt in T]
            rec["predict_mu"] = [f17(p[0]) for p in preds]
            rec["predict_var"] = [f17(p[1]) for p in preds]
            zi = update_w(root, infer=True)
            rec["infer_z"] = f17(zi)
            rec["infer_logw"] = [[f17(v) for v in n.logw] if n.type >= SUM else [] for n in order]
        else:
            K = len(root.children)
            poe_mu, poe_var, g_mu, g_var, r_mu, r_var = [], [], [], [], [], []
            for xt in T:
                m, t = poe_rec(root, xt)                                   # predictPoE :256-260
                poe_mu.append(f17(m)); poe_var.append(f17(1 / t))
                beta = mp.mpf(1) / K                                       # _predictgPoE :211-222
                tt = mp.mpf(0); mm = mp.mpf(0)
                for c in root.children:
                    m_, t_ = poe_rec(c, xt)
                    tt += beta * t_; mm += beta * t_ * m_
                g_mu.append(f17(mm / tt)); g_var.append(f17(1 / tt))
                gp0 = leftgp(root)                                          # _predictrBCM :224-241
                s = kval(gp0.ktype, gp0.th, xt, xt) + gp0.eta
                C = 1 / s; mu = mp.mpf(0)
                for c in root.children:
                    m_, t_ = poe_rec(c, xt)
                    s_ = 1 / t_
                    b_ = mp.mpf("0.5") * (mp.log(s) - mp.log(s_))
                    C += (b_ * t_) - (b_ / s)
                    mu += m_ * (b_ * t_)
                r_mu.append(f17(mu / C)); r_var.append(f17(1 / C))
            rec.update(poe_mu=poe_mu, poe_var=poe_var, gpoe_mu=g_mu, gpoe_var=g_var, rbcm_mu=r_mu, rbcm_var=r_var)
        # single-expert predictions of leaf 0 on all test points (prediction(gp, x))
        rec["leaf0_mu"] = [f17(leaves[0].gp.predict(xt)[0]) for xt in T]
        rec["leaf0_var"] = [f17(leaves[0].gp.predict(xt)[1]) for xt in T]
        out["evals"].append(rec)
    return out


def main():
    X, y, T = make_data()
    models = []
    for nm, kt in (("isose", ISO_SE), ("ardse", ARD_SE), ("isolinear", ISO_LINEAR), ("ardlinear", ARD_LINEAR)):
        models.append(run_model("dsmgp_" + nm, X, y, T, [kt]))
    models.append(run_model("dsmgp_mixture", X, y, T, [ISO_SE, ISO_LINEAR]))
    models.append(run_model("poe_isose", X, y, T, [ISO_SE], poe=True))
    doc = dict(generator="tests/golden/make_mp_golden.py (mpmath %s, %d digits; does not import the oracle)" % (mp.__version__, mp.mp.dps),
               x=X, y=y, xtest=T, models=models)
    with open(os.path.join(HERE, "mp_golden.json"), "w") as f:
        json.dump(doc, f)
    print("wrote mp_golden.json:", [m["name"] for m in models])


if __name__ == "__main__":
    main()
