"""The oracle restatement pinned against itself: finite differences, mpmath, LAPACK identities, the reference's
in-source self tests restated (AdvancedCholeskey.jl:61-135), and the committed golden vectors.
(The reference ships no tests or fixtures -- SURVEY §4 -- so these are the pin.)"""
import json
import math
import os

import numpy as np
import pytest
import scipy.linalg as sla

from conftest import orc, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")

KERNELS = {
    "isose": lambda D: orc.IsoSE(0.1, -0.2),
    "ardse": lambda D: orc.ArdSE(np.linspace(-0.3, 0.2, D), -0.2),
    "isolin": lambda D: orc.IsoLinear(0.2),
    "ardlin": lambda D: orc.ArdLinear(np.linspace(-0.1, 0.2, D)),
}


def fd_grad(x, y, k, logn, h=1e-6):
    th0 = np.concatenate([k.logl, [k.logs, logn]])

    def f(th):
        g = orc.GaussianProcess(x, y, kernel=k, logNoise=logn)
        g.setparams(th)
        return g.update_cholesky().mll()

    return np.array([(f(th0 + h * e) - f(th0 - h * e)) / (2 * h) for e in np.eye(th0.size)])


@pytest.mark.parametrize("name", list(KERNELS))
def test_mathematical_gradient_matches_finite_differences(name):
    x, y = synth(40, 3, 1)
    k = KERNELS[name](3)
    gp = orc.GaussianProcess(x, y, kernel=k, logNoise=-0.5, run_cholesky=True)
    fd = fd_grad(x, y, k, -0.5)
    g = gp.grad_mll(mathematical=True)
    assert np.allclose(g, fd, rtol=2e-6, atol=2e-6)


def test_as_written_quirks_q2_q3():
    """Q2 extra exp(log sigma) factor on kernel gradients; Q3 ArdSE length-scale gradient == 0;
    the noise gradient is exact."""
    x, y = synth(40, 3, 2)
    s = math.exp(-0.2)
    k = KERNELS["isose"](3)
    gp = orc.GaussianProcess(x, y, kernel=k, logNoise=-0.5, run_cholesky=True)
    fd = fd_grad(x, y, k, -0.5)
    g = gp.grad_mll(as_written_dense=True)
    assert np.allclose(g[:2], s * fd[:2], rtol=2e-6) and np.isclose(g[2], fd[2], rtol=2e-6)
    k = KERNELS["ardse"](3)
    gp = orc.GaussianProcess(x, y, kernel=k, logNoise=-0.5, run_cholesky=True)
    fd = fd_grad(x, y, k, -0.5)
    g = gp.grad_mll(as_written_dense=True)
    assert np.all(g[:3] == 0.0) and np.all(np.abs(fd[:3]) > 1e-3)
    assert np.isclose(g[3], s * fd[3], rtol=2e-6) and np.isclose(g[4], fd[4], rtol=2e-6)


@pytest.mark.parametrize("name", list(KERNELS))
def test_fast_gradient_equals_dense_as_written(name):
    x, y = synth(60, 3, 3)
    k = KERNELS[name](3)
    gp = orc.GaussianProcess(x, y, kernel=k, logNoise=-0.7, run_cholesky=True)
    a = gp.grad_mll(as_written_dense=True).copy()
    b = gp.grad_mll().copy()
    assert np.allclose(a, b, rtol=1e-10, atol=1e-11)


def test_ardse_is_additive_q1():
    x, _ = synth(10, 4, 4)
    k = orc.ArdSE(np.zeros(4), 0.3)
    K = orc.kernelmatrix(k, x)
    assert np.allclose(np.diag(K), 4 * math.exp(0.6))
    ref = sum(np.exp(-0.5 * (x[:, None, d] - x[None, :, d]) ** 2) for d in range(4)) * math.exp(0.6)
    assert np.allclose(K, ref, rtol=1e-15)


def test_lml_against_mpmath():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    x, y = synth(12, 2, 5)
    k = orc.IsoSE(0.3, 0.1)
    gp = orc.GaussianProcess(x, y, kernel=k, logNoise=-0.4, run_cholesky=True)
    n = 12
    F = mp.matrix(n, n)
    l2 = mp.e ** (2 * mp.mpf(0.3)); v = mp.e ** (2 * mp.mpf(0.1)); c = mp.e ** (2 * mp.mpf(-0.4)) + mp.mpf("1e-8")
    for i in range(n):
        for j in range(n):
            r2 = sum((mp.mpf(float(x[i, d])) - mp.mpf(float(x[j, d]))) ** 2 for d in range(2))
            F[i, j] = v * mp.e ** (-r2 / (2 * l2)) + (c if i == j else 0)
    yc = mp.matrix([mp.mpf(float(t)) for t in gp.y])
    alpha = mp.lu_solve(F, yc)
    lml = -((yc.T * alpha)[0] + mp.log(mp.det(F)) + n * mp.log(2 * mp.pi)) / 2
    assert abs(float(lml) - gp.mll()) < 1e-11 * abs(float(lml))


def test_chol_continue_restated():
    """test_chol_continue AdvancedCholeskey.jl:121-135."""
    rng = np.random.default_rng(6)
    S = orc.gen_cov(100, rng)
    A = S.copy()
    A[:10, :10] = sla.cholesky(S[:10, :10], lower=True)
    out, info = orc.chol_continue(A, 11)
    assert info == 0
    assert np.sum(np.abs(out - sla.cholesky(S, lower=True))) < 1e-10


def test_lrtest_restated_corrected_and_as_written():
    """lrtest AdvancedCholeskey.jl:61-110: the corrected row deletion matches a fresh factorisation; the loop as
    written (App. B Q7) does not."""
    rng = np.random.default_rng(7)
    D = 60
    S = orc.gen_cov(D, rng)
    missing = np.sort(rng.permutation(D - 1)[:5])
    keep = np.setdiff1d(np.arange(D), missing)
    Lf = sla.cholesky(S, lower=True)
    ref = sla.cholesky(S[np.ix_(keep, keep)], lower=True)
    good = orc.chol_delete_rows(Lf, missing.tolist())
    assert np.sum(np.abs(good - ref)) < 1e-11
    CC = Lf.copy()
    for r in missing:       # 0-based r -> reference call lowrankupdate!(CC, view(CC, r+1:D, r), r+1) with 1-based r
        CC = orc.lowrankupdate_as_written(CC, CC[r + 1:, r].copy(), r + 2)
    bad = np.tril(CC)[np.ix_(keep, keep)]
    assert np.sum(np.abs(bad - ref)) > 1e-3


def test_prediction_matches_direct_formula():
    x, y = synth(50, 2, 8)
    k = orc.ArdSE([0.1, -0.1], 0.2)
    gp = orc.GaussianProcess(x, y, kernel=k, logNoise=-0.6, run_cholesky=True)
    xt = np.random.default_rng(1).random((7, 2))
    mu, s2 = gp.prediction(xt)
    mu2, S = gp.prediction(xt, full_cov=True)
    F = orc.kernelmatrix(k, x) + (gp.noise() + 1e-8) * np.eye(50)
    Knt = orc.kernelmatrix(k, x, xt)
    assert np.allclose(mu, gp.mean + Knt.T @ np.linalg.solve(F, gp.y), rtol=1e-10)
    assert np.allclose(s2, np.diag(S), rtol=1e-12)
    assert np.allclose(np.diag(S), np.diag(orc.kernelmatrix(k, xt) - Knt.T @ np.linalg.solve(F, Knt)) + gp.noise(), rtol=1e-9)


def test_golden_vectors():
    """Committed outputs of the oracle (tests/golden/make_golden.py) guard against silent drift of the checker."""
    from golden.make_golden import CASES, run_case
    with open(os.path.join(GOLD, "golden.json")) as f:
        gold = json.load(f)
    for name in CASES:
        out = run_case(name)
        g = gold[name]
        assert abs(out["lml"] - g["lml"]) <= 1e-11 * abs(g["lml"]), name
        assert np.allclose(out["grad"], g["grad"], rtol=1e-9, atol=1e-9), name
        assert np.allclose(out["mu"], g["mu"], rtol=1e-10) and np.allclose(out["var"], g["var"], rtol=1e-9), name


def test_table_exp_restated_accuracy():
    """csrc/fastexp.cuh (`exp_neg`: 64-entry 2^(j/64) table + degree-5 polynomial) restated in NumPy, against mpmath:
    the Gram / predict / LAUUM kernels use it instead of the library exp(), so its error budget is part of parity."""
    import re, os
    import mpmath as mp
    mp.mp.dps = 40
    src = open(os.path.join(os.path.dirname(__file__), "..", "deepstructuredmixtures_b200", "csrc", "fastexp.cuh")).read()
    body = src[src.index("g_exptab[EXPTAB_N] = {") + len("g_exptab[EXPTAB_N] = {"):src.index("};")]
    T = np.array([float(v) for v in re.findall(r"[0-9.eE+-]+", body)])
    assert T.size == 64 and all(T[j] == float(mp.mpf(2) ** (mp.mpf(j) / 64)) for j in range(64))
    MAGIC = 6755399441055744.0

    def exp_neg(x):
        t = x * 92.33248261689366 + MAGIC
        k = (t.view(np.int64) & 0xFFFFFFFF).astype(np.int64)
        k = np.where(k >= 2 ** 31, k - 2 ** 32, k)
        kf = t - MAGIC
        r = x - kf * 0.01083042469326756
        r = r - kf * 2.9815858269852933e-12
        p = r * 8.3333333333333332e-03 + 4.1666666666666664e-02
        p = r * p + 1.6666666666666666e-01
        p = r * p + 0.5
        p = r * p + 1.0
        p = p * r
        v = T[k & 63] * p + T[k & 63]
        return np.where(x < -708.0, 0.0, np.ldexp(v, (k >> 6).astype(np.int32)))

    rng = np.random.default_rng(0)
    x = -np.abs(rng.standard_normal(4000)) * np.array([1e-3, 1.0, 30.0, 200.0])[rng.integers(0, 4, 4000)]
    x = np.concatenate([x[x > -700.0], [0.0, -1e-300, -707.9]])
    ref = np.array([float(mp.e ** mp.mpf(float(v))) for v in x])
    assert np.max(np.abs(exp_neg(x) - ref) / ref) < 4.5e-16
    assert exp_neg(np.array([-709.0]))[0] == 0.0 and exp_neg(np.array([0.0]))[0] == 1.0


# ---- the independent pin: 50-digit mpmath vectors that do not share code with the oracle -------------------------------
def _mp_models():
    with open(os.path.join(GOLD, "mp_golden.json")) as f:
        doc = json.load(f)
    return doc, {m["name"]: m for m in doc["models"]}


def mp_oracle_root(doc, model, theta):
    flat = {k: np.asarray(v) for k, v in model["flat"].items() if k != "root"}
    flat["root"] = model["flat"]["root"]
    D = len(doc["x"][0])
    kernels = []
    for kt in model["kernels"]:
        nl = 1 if kt in (orc.ISO_SE, orc.ISO_LINEAR) else D
        kernels.append(orc.Kernel(kt, np.zeros(nl), 0.0))
    root = orc.tree_from_flat(flat, np.asarray(doc["x"]), np.asarray(doc["y"]), kernels, -1.0)
    return root


MP_NAMES = ["dsmgp_isose", "dsmgp_ardse", "dsmgp_isolinear", "dsmgp_ardlinear", "dsmgp_mixture", "poe_isose"]


@pytest.mark.parametrize("name", MP_NAMES)
def test_oracle_against_independent_mpmath_vectors(name):
    """tests/golden/mp_golden.json is written by make_mp_golden.py, which evaluates the Julia formulas in mpmath at 50 digits
    and imports nothing from oracle/: LML, as-written and mathematical gradients, mll!, the down-pass (plain and finetune),
    update!, infer!, DSMGP / PoE / gPoE / rBCM predictions of the oracle must agree with it."""
    doc, models = _mp_models()
    m = models[name]
    xt = np.asarray(doc["xtest"])
    worst = 0.0
    for ev in m["evals"]:
        theta = np.asarray(ev["theta"])
        root = mp_oracle_root(doc, m, theta)
        for mode, math_ in (("as_written", False), ("mathematical", True)):
            lml, grad, ell, rows = orc.evaluate(root, theta, mathematical=math_, as_written_dense=not math_ and "ardlinear" not in name)
            assert abs(lml - ev["node_lml"][m["flat"]["root"]]) <= 1e-12 * abs(lml)
            for nid, v in ell.items():
                assert abs(v - ev["node_lml"][nid]) <= 1e-12 * max(1.0, abs(v))
            for l, r in rows.items():
                g = np.asarray(ev["leaf_grad_" + mode][l])
                scale = np.maximum(np.abs(g), 1e-10 * max(1.0, np.max(np.abs(g))))
                err = np.max(np.abs(r[1:] - g) / np.maximum(scale, 1e-3))
                worst = max(worst, err)
                assert err <= 1e-9, (name, mode, l, r[1:], g)
                assert abs(r[0] - ev["leaf_lml"][l]) <= 1e-12 * abs(r[0])
            G = np.asarray(ev["grad_" + mode])
            assert np.all(np.abs(grad - G) <= 1e-9 * np.maximum(np.abs(G), 1e-3)), (name, mode, grad, G)
            _, gradf, _, _ = orc.evaluate(root, theta, Drow=np.asarray(ev["finetune_row"]), mathematical=math_)
            Gf = np.asarray(ev["grad_" + mode + "_finetune"])
            assert np.all(np.abs(gradf - Gf) <= 1e-9 * np.maximum(np.abs(Gf), 1e-3))
        orc.evaluate(root, theta)
        mu0, var0 = orc.getLeaves(root)[0].gp.prediction(xt)
        assert np.allclose(mu0, ev["leaf0_mu"], rtol=1e-10, atol=1e-12) and np.allclose(var0, ev["leaf0_var"], rtol=1e-10)
        if name.startswith("dsmgp"):
            z = orc.update_weights(root)
            assert abs(z - ev["update_z"]) <= 1e-12 * abs(z)
            for nid, lw in enumerate(ev["update_logw"]):
                if lw:
                    node = _node_by_id(root, nid)
                    assert np.allclose(node.logweights, lw, rtol=1e-10, atol=1e-12)
            mu, var = orc.predict_dsmgp(root, xt)
            assert np.allclose(mu, ev["predict_mu"], rtol=1e-9, atol=1e-11) and np.allclose(var, ev["predict_var"], rtol=1e-9)
            zi = orc.infer_weights(root)
            assert abs(zi - ev["infer_z"]) <= 1e-12 * abs(zi)
            for nid, lw in enumerate(ev["infer_logw"]):
                if lw:
                    assert np.allclose(_node_by_id(root, nid).logweights, lw, rtol=1e-10, atol=1e-12)
        else:
            for fn, key in ((orc.predict_poe, "poe"), (orc.predict_gpoe, "gpoe"), (orc.predict_rbcm, "rbcm")):
                mu, var = fn(root, xt)
                assert np.allclose(mu, ev[key + "_mu"], rtol=1e-9, atol=1e-11) and np.allclose(var, ev[key + "_var"], rtol=1e-9), key
    print(f"{name}: worst leaf-gradient error vs mpmath {worst:.2e}")


def _node_by_id(root, nid):
    if root.id == nid:
        return root
    for c in root.children:
        r = _node_by_id(c, nid)
        if r is not None:
            return r
    return None
