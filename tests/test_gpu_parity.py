"""GPU parity: libdsmgp (through the C ABI via ctypes) against the CPU oracle on identical seeded inputs.
Tolerances are those of BASELINE.json north_star: LML and gradients 1e-9 relative, predictions 1e-8 relative
(gradient components that are differences of large terms are compared relative to the larger term)."""
import math

import numpy as np
import pytest

from conftest import oracle_tree, orc, relerr, synth, theta0

pytestmark = pytest.mark.gpu

LML_TOL = 1e-9
GRAD_TOL = 1e-9
PRED_TOL = 1e-8


def pred_close(mu, omu, var, ovar, yscale):
    """Predictions within PRED_TOL relative.  The mean m + K' alpha is a sum of large cancelling terms whose rounding
    error scales with the data, not with the (possibly ~0) value, so |mu| is floored by the scale of the targets."""
    ok_mu = np.all(np.abs(mu - omu) <= PRED_TOL * np.maximum(np.abs(omu), yscale))
    ok_var = np.all(np.abs(var - ovar) <= PRED_TOL * np.abs(ovar))
    return bool(ok_mu and ok_var)


def check_eval(model, theta, mathematical=False, leaf_scale=None):
    import deepstructuredmixtures_b200 as dsm
    lml, grad, nodes = model.handle.eval(theta, leaf_scale=leaf_scale, want_nodes=True)
    rows = model.handle.leaf_rows()
    root = oracle_tree(model)
    o_lml, o_grad, o_ell, o_rows = orc.evaluate(root, theta, Drow=leaf_scale, mathematical=mathematical)
    assert abs(lml - o_lml) <= LML_TOL * abs(o_lml), (lml, o_lml)
    # per-node table
    for nid, v in o_ell.items():
        assert abs(nodes[nid] - v) <= LML_TOL * max(abs(v), 1.0)
    # per-leaf rows: LML tight; gradients relative to the natural scale of their two large terms
    for l, orow in o_rows.items():
        assert abs(rows[l, 0] - orow[0]) <= LML_TOL * abs(orow[0]), (l, rows[l, 0], orow[0])
        gp = [lf for lf in orc.getLeaves(root) if lf.leaf_index == l][0].gp
        scale = max(float(gp.alpha @ gp.alpha), float(gp.N)) * max(1.0, gp.noise())
        g = rows[l, 1:1 + orow.size - 1]
        if leaf_scale is not None and leaf_scale[l] == 0.0:
            assert np.all(g == 0.0)          # zero-weight experts skip the gradient kernels (include/dsmgp.h: dsmgp_finetune_eval)
            continue
        # 1e-9 RELATIVE; a component that is smaller than 1e-4 of its two cancelling terms is compared with that floor
        assert np.all(np.abs(g - orow[1:]) <= GRAD_TOL * np.maximum(np.abs(orow[1:]), 1e-4 * scale)), (l, g, orow[1:])
    # model gradient (optimize.jl:42-89: sum_l w_l g_l): 1e-9 relative, where a component that is a cancelling sum of leaf
    # gradients is compared with the natural scale sum_l w_l scale_l of its terms (the down-pass applied to the leaf scales)
    leaves = orc.getLeaves(root)
    nat_scale = np.zeros_like(o_grad)
    orc.grad_down(root, 0.0, 0.0, o_ell, o_ell[root.id], nat_scale,
                  {lf.leaf_index: np.full(lf.gp.nparams(), max(float(lf.gp.alpha @ lf.gp.alpha), float(lf.gp.N)) * max(1.0, lf.gp.noise()))
                   for lf in leaves}, leaf_scale)
    err_scaled = np.max(np.abs(grad - o_grad) / np.maximum(np.abs(o_grad), nat_scale))
    nz = np.abs(o_grad) > 0
    err_rel = np.max(np.abs(grad - o_grad)[nz] / np.abs(o_grad)[nz]) if nz.any() else 0.0
    print(f"\n[parity] L={len(leaves)} lml rel {abs(lml - o_lml) / abs(o_lml):.1e}  model grad rel {err_rel:.1e} (rel to natural scale {err_scaled:.1e})")
    assert np.all(np.abs(grad - o_grad) <= GRAD_TOL * np.maximum(np.abs(o_grad), 1e-4 * nat_scale)), (grad, o_grad)
    return lml, grad


def test_single_gp_isose():
    import deepstructuredmixtures_b200 as dsm
    x, y = synth(300, 2, 11)
    gp = dsm.GaussianProcess(x, y, kernel=dsm.IsoSE(0.1, -0.2), logNoise=-1.0, run_cholesky=True)
    o = orc.GaussianProcess(x, y, kernel=orc.IsoSE(0.1, -0.2), logNoise=-1.0, run_cholesky=True)
    assert abs(gp.mll() - o.mll()) <= LML_TOL * abs(o.mll())
    assert np.max(np.abs(gp.alpha - o.alpha)) <= 1e-9 * np.max(np.abs(o.alpha))      # norm-wise: entries of alpha cross zero
    Lf = gp.factors
    assert np.max(np.abs(Lf - np.tril(o.L))) < 1e-11
    g = dsm.grad_mll(gp)
    og = o.grad_mll(as_written_dense=True)
    assert np.all(np.abs(g - og) <= GRAD_TOL * np.maximum(np.abs(og), 1e-4 * float(o.N)))
    xt = np.random.default_rng(5).random((77, 2))
    mu, var = gp.prediction(xt)
    omu, ovar = o.prediction(xt)
    assert pred_close(mu, omu, var, ovar, float(np.std(y)))


@pytest.mark.parametrize("ktype", ["isose", "ardse", "isolin", "ardlin"])
@pytest.mark.parametrize("n", [5, 64, 65, 129, 200, 333])
def test_single_gp_sizes(ktype, n):
    import deepstructuredmixtures_b200 as dsm
    D = 3
    x, y = synth(n, D, 100 + n)
    kern = {"isose": dsm.IsoSE(0.2, 0.1), "ardse": dsm.ArdSE([0.1, -0.3, 0.4], 0.2),
            "isolin": dsm.IsoLinear(0.3), "ardlin": dsm.ArdLinear([0.2, -0.1, 0.5])}[ktype]
    okern = orc.Kernel(kern.type, kern.logl, kern.logs)
    gp = dsm.GaussianProcess(x, y, kernel=kern, logNoise=-0.7, run_cholesky=True)
    o = orc.GaussianProcess(x, y, kernel=okern, logNoise=-0.7, run_cholesky=True)
    assert abs(gp.mll() - o.mll()) <= LML_TOL * abs(o.mll())
    g = dsm.grad_mll(gp)
    og = o.grad_mll(as_written_dense=(ktype != "ardlin"))
    scale = max(float(o.alpha @ o.alpha), float(n))
    assert np.all(np.abs(g - og) <= GRAD_TOL * np.maximum(np.abs(og), 1e-4 * scale)), (g, og)
    xt = np.random.default_rng(n).random((50, D))
    mu, var = gp.prediction(xt)
    omu, ovar = o.prediction(xt)
    assert pred_close(mu, omu, var, ovar, float(np.std(y)))


def test_cfg1_readme_dsmgp():
    """BASELINE configs[0]: README example, 1-D N=100, IsoSE(1,1) + ConstMean, V=3 K=4 M=10."""
    import deepstructuredmixtures_b200 as dsm
    rng = np.random.default_rng(1)
    x = np.linspace(0, 1, 100)
    y = np.sin(x * 4 * np.pi + rng.standard_normal(100) * 0.2)
    model = dsm.buildDSMGP(x.reshape(-1, 1), y, 3, 4, M=10, kernel=dsm.IsoSE(1.0, 1.0),
                           meanFun=dsm.ConstMean(float(np.mean(x))), rng=1)
    th = np.array([-1.0, 0.2, -1.2])
    check_eval(model, th)
    z = dsm.update_(model)
    root = oracle_tree(model, th)
    orc.fit(root)
    oz = orc.update_weights(root)
    assert abs(z - oz) <= LML_TOL * abs(oz)
    xt = np.linspace(0.0, 1.0, 173).reshape(-1, 1)
    mu, var = dsm.predict(model, xt)
    omu, ovar = orc.predict_dsmgp(root, xt)
    assert pred_close(mu, omu, var, ovar, float(np.std(y)))


@pytest.mark.parametrize("mathematical", [False, True])
def test_dsmgp_ardse_small(mathematical):
    import deepstructuredmixtures_b200 as dsm
    x, y = synth(1500, 4, 3)
    model = dsm.buildDSMGP(x, y, 2, 3, M=60, kernel=dsm.ArdSE(np.zeros(4), 0.0), logNoise=-1.0, rng=3,
                           as_written_grads=not mathematical)
    th = np.array([0.1, -0.2, 0.3, 0.0, 0.2, -1.0])
    check_eval(model, th, mathematical=mathematical)


def test_dsmgp_kernel_mixture():
    import deepstructuredmixtures_b200 as dsm
    x, y = synth(1200, 3, 4)
    model = dsm.buildDSMGP(x, y, 2, 2, M=100, kernel=[dsm.IsoSE(0.0, 0.0), dsm.IsoLinear(0.0)], logNoise=-1.0, rng=4)
    th = np.array([0.2, 0.1, -1.0, 0.3, 0.0, -0.8])
    check_eval(model, th)


def test_finetune_leaf_scale():
    import deepstructuredmixtures_b200 as dsm
    x, y = synth(900, 2, 5)
    model = dsm.buildDSMGP(x, y, 2, 3, M=50, kernel=dsm.IsoSE(0.0, 0.0), logNoise=-1.0, rng=5)
    D = model.Dmat if hasattr(model, "Dmat") else model.D
    root = oracle_tree(model)
    oD = orc.getOverlap(root, x.shape[0])
    assert np.array_equal(D, oD)
    th = np.array([-0.5, 0.1, -1.1])
    check_eval(model, th, leaf_scale=D[1, :])


@pytest.mark.parametrize("kernel", ["isose", "mixture"])
def test_overlap_matrix_device_bit_exact(kernel):
    """dsmgp_overlap == getOverlap (fit.jl:12-39) bit for bit: oracle restatement and the host restatement."""
    import deepstructuredmixtures_b200 as dsm
    from deepstructuredmixtures_b200 import structure as st
    x, y = synth(1500, 3, 8)
    k = dsm.IsoSE(0.0, 0.0) if kernel == "isose" else [dsm.IsoSE(0.0, 0.0), dsm.IsoLinear(0.0)]
    model = dsm.buildDSMGP(x, y, 3, 3, M=60, kernel=k, logNoise=-1.0, rng=8)
    D = model.D
    assert np.array_equal(D, orc.getOverlap(oracle_tree(model), x.shape[0]))
    assert np.array_equal(D, st.getOverlap(model.root, x.shape[0]))
    assert D.max() <= 1.0 and D.min() >= 0.0 and np.all(np.diag(D) == 0.0) and np.count_nonzero(D) > 0
    # the sparse form (dsmgp_overlap_csr): exactly the non-zeros of D, row by row, columns ascending
    from deepstructuredmixtures_b200.linalg import overlap_matrix_csr
    rp, col, val = overlap_matrix_csr(x.shape[0], [lf.obs for lf in model.leaves], [lf.kernelid - 1 for lf in model.leaves], model.flat)
    assert rp[-1] == np.count_nonzero(D) and rp[-1] < D.size
    S = np.zeros_like(D)
    for n in range(D.shape[0]):
        c = col[rp[n]:rp[n + 1]]
        assert np.all(np.diff(c) > 0)
        S[n, c] = val[rp[n]:rp[n + 1]]
    assert np.array_equal(S, D)


@pytest.mark.parametrize("ktype,mathematical", [("isose", False), ("ardse", True), ("ardlin", False)])
def test_wide_inputs_borrowed_ring(ktype, mathematical):
    """D = 12 > 8: the LAUUM epilogue and the predict kernel stage their point tiles in the (borrowed) pipeline ring
    instead of the static shared buffer; multi-block experts so that the producer hand-shake is exercised."""
    import deepstructuredmixtures_b200 as dsm
    D = 12
    x, y = synth(900, D, 21)
    k = {"isose": dsm.IsoSE(0.3, 0.1), "ardse": dsm.ArdSE(np.linspace(0.2, 0.5, D), 0.1),
         "ardlin": dsm.ArdLinear(np.linspace(0.1, 0.4, D))}[ktype]
    model = dsm.buildDSMGP(x, y, 2, 2, M=150, kernel=k, logNoise=-1.0, rng=21, as_written_grads=not mathematical)
    th = model.handle.get_leaf_params(0).copy()
    check_eval(model, th, mathematical=mathematical)
    dsm.fit_(model)
    dsm.update_(model)
    xt = np.random.default_rng(9).random((300, D))
    mu, var = dsm.predict(model, xt)
    root = oracle_tree(model)
    orc.setparams(root, th); orc.fit(root); orc.update_weights(root)
    omu, ovar = orc.predict_dsmgp(root, xt)
    assert pred_close(mu, omu, var, ovar, float(np.std(y)))


def test_predict_task_granularities_agree(monkeypatch):
    """predict3 runs one task per (expert, 128 test points) when there are many test points and one task per
    (expert, 128 test points, row block) with cross-CTA flags when there are few: both against the oracle."""
    import deepstructuredmixtures_b200 as dsm
    x, y = synth(2500, 3, 31)
    model = dsm.buildDSMGP(x, y, 2, 2, M=300, kernel=dsm.ArdSE(np.zeros(3), 0.0), logNoise=-1.0, rng=31)
    th = np.array([0.1, -0.1, 0.2, 0.1, -1.0])
    model.handle.eval(th)
    dsm.update_(model)
    xt = np.random.default_rng(4).random((700, 3))
    root = oracle_tree(model, th)
    orc.fit(root); orc.update_weights(root)
    omu, ovar = orc.predict_dsmgp(root, xt)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("DSMGP_PREDICT_WAVE", mode)
        mu, var = dsm.predict(model, xt)
        assert pred_close(mu, omu, var, ovar, float(np.std(y))), mode
        out[mode] = (mu, var)
    assert np.max(np.abs(out["0"][0] - out["1"][0])) <= 1e-10 * np.max(np.abs(omu))
    assert np.max(np.abs(out["0"][1] - out["1"][1]) / np.abs(ovar)) <= 1e-10


def test_finetune_eval_batched():
    """dsmgp_finetune_eval (one call for all anchors) == finetuning.jl:36-58 evaluated anchor by anchor by the oracle."""
    import deepstructuredmixtures_b200 as dsm
    x, y = synth(700, 2, 6)
    model = dsm.buildDSMGP(x, y, 2, 3, M=40, kernel=dsm.IsoSE(0.0, 0.0), logNoise=-1.0, rng=6)
    D = model.D
    L = len(model.leaves)
    rng = np.random.default_rng(3)
    thetas = np.array([-0.4, 0.1, -1.0]) + 0.2 * rng.standard_normal((L, 3))
    leaf_lml, grads, root_lml = model.handle.finetune_eval(np.arange(L), thetas, D)
    root = oracle_tree(model)
    node_of_leaf = {lf.leaf_index: lf.id for lf in orc.getLeaves(root)}
    for g in range(L):
        o_lml, o_grad, o_ell, _ = orc.evaluate(root, thetas[g], Drow=D[g, :])
        assert abs(root_lml[g] - o_lml) <= LML_TOL * abs(o_lml)
        assert abs(leaf_lml[g] - o_ell[node_of_leaf[g]]) <= LML_TOL * abs(o_ell[node_of_leaf[g]])
        gs = np.maximum(np.abs(o_grad), 1e-6 * np.max(np.abs(o_grad)) + 1e-300)
        assert np.all(np.abs(grads[g] - o_grad) <= 1e-8 * np.maximum(gs, 1.0)), (g, grads[g], o_grad)
    # and the driver: two finetune_ iterations run and keep per-leaf parameters
    m2, ell = dsm.finetune_(model, dsm.ADAM(), iterations=2)
    assert np.all(np.isfinite(ell))


@pytest.mark.parametrize("mode", ["poe", "gpoe", "rbcm"])
def test_poe_family_predict(mode):
    import deepstructuredmixtures_b200 as dsm
    x, y = synth(800, 2, 6)
    kw = dict(M=60, kernel=dsm.IsoSE(-0.5, 0.0), logNoise=-1.0, rng=6)
    model = dsm.buildBCM(x, y, 3, **kw) if mode == "rbcm" else dsm.buildPoE(x, y, 3, generalized=(mode == "gpoe"), **kw)
    root = oracle_tree(model)
    orc.fit(root)
    xt = np.random.default_rng(9).random((120, 2))
    mu, var = dsm.predict(model, xt)
    f = {"poe": orc.predict_poe, "gpoe": orc.predict_gpoe, "rbcm": orc.predict_rbcm}[mode]
    omu, ovar = f(root, xt)
    assert pred_close(mu, omu, var, ovar, float(np.std(y)))


def test_kernelmatrix_operator():
    import deepstructuredmixtures_b200 as dsm
    rng = np.random.default_rng(7)
    x1, x2 = rng.random((150, 5)), rng.random((97, 5))
    for k, ok in [(dsm.IsoSE(0.3, -0.1), orc.IsoSE(0.3, -0.1)),
                  (dsm.ArdSE(np.linspace(-0.2, 0.4, 5), 0.2), orc.ArdSE(np.linspace(-0.2, 0.4, 5), 0.2)),
                  (dsm.IsoLinear(0.1), orc.IsoLinear(0.1)),
                  (dsm.ArdLinear(np.linspace(-0.2, 0.4, 5)), orc.ArdLinear(np.linspace(-0.2, 0.4, 5)))]:
        K = dsm.kernelmatrix(k, x1, x2)
        assert np.max(np.abs(K - orc.kernelmatrix(ok, x1, x2))) < 1e-13 * max(1.0, np.max(np.abs(K)))


def test_chol_continue_restated():
    """AdvancedCholesky.test_chol_continue (AdvancedCholeskey.jl:121-135): D = 100, P = 10."""
    import deepstructuredmixtures_b200 as dsm
    import scipy.linalg as sla
    rng = np.random.default_rng(8)
    for D, P in [(100, 10), (300, 128), (257, 130), (64, 63)]:
        S = orc.gen_cov(D, rng)
        A = S.copy()
        A[:P, :P] = np.tril(sla.cholesky(S[:P, :P], lower=True)) + np.triu(S[:P, :P], 1)
        out, info = dsm.chol_continue_(A, P + 1)
        ref = sla.cholesky(S, lower=True)
        assert info == 0
        assert np.sum(np.abs(out - ref)) < 1e-9
        o2, oinfo = orc.chol_continue(A, P + 1)
        assert np.max(np.abs(out - o2)) < 1e-11


def test_chol_delete_rows_lrtest():
    """AdvancedCholesky.lrtest (AdvancedCholeskey.jl:61-110) with the corrected algorithm."""
    import deepstructuredmixtures_b200 as dsm
    import scipy.linalg as sla
    rng = np.random.default_rng(9)
    D = 200
    S = orc.gen_cov(D, rng)
    missing = np.sort(rng.permutation(D - 1)[:10])
    keep = np.setdiff1d(np.arange(D), missing)
    Lf = sla.cholesky(S, lower=True)
    out = dsm.chol_delete_rows(Lf, (missing + 1).tolist())
    ref = sla.cholesky(S[np.ix_(keep, keep)], lower=True)
    assert np.sum(np.abs(out - ref)) < 1e-9
    assert np.max(np.abs(out - orc.chol_delete_rows(Lf, missing.tolist()))) < 1e-11
    # batched form: several factors, one launch, one column sweep per matrix for all of its rows (incl. > 64 rows: two chunks)
    from deepstructuredmixtures_b200.linalg import chol_delete_rows_batched
    mats, rws, refs = [], [], []
    for D2, nd in ((150, 3), (333, 70), (64, 1), (200, 10)):
        S2 = orc.gen_cov(D2, rng)
        miss = np.sort(rng.permutation(D2 - 1)[:nd])
        kp = np.setdiff1d(np.arange(D2), miss)
        mats.append(sla.cholesky(S2, lower=True)); rws.append((miss + 1).tolist())
        refs.append(sla.cholesky(S2[np.ix_(kp, kp)], lower=True))
    outs = chol_delete_rows_batched(mats, rws)
    for o, r_ in zip(outs, refs):
        assert np.max(np.abs(o - r_)) < 1e-10 * np.max(np.abs(r_))
    assert np.array_equal(outs[3], dsm.chol_delete_rows(mats[3], rws[3]))


def test_not_positive_definite_reports_info():
    import deepstructuredmixtures_b200 as dsm
    A = np.eye(70); A[40, 40] = -1.0
    _, info = dsm.potrf_(A)
    assert info == 41


@pytest.mark.parametrize("n,ktype", [(1500, "ardse"), (1088, "isose"), (1345, "isolin"), (2100, "ardse")])
def test_single_gp_multiblock(n, ktype):
    """Many 128-blocks per expert (nb = 9..17), with and without a trailing half tile."""
    import deepstructuredmixtures_b200 as dsm
    D = 4
    x, y = synth(n, D, 7 + n)
    kern = {"isose": dsm.IsoSE(-0.5, 0.1), "ardse": dsm.ArdSE([-0.5, -0.3, -0.4, -0.6], 0.2), "isolin": dsm.IsoLinear(0.3)}[ktype]
    okern = orc.Kernel(kern.type, kern.logl, kern.logs)
    gp = dsm.GaussianProcess(x, y, kernel=kern, logNoise=-1.0, run_cholesky=True)
    o = orc.GaussianProcess(x, y, kernel=okern, logNoise=-1.0, run_cholesky=True)
    assert abs(gp.mll() - o.mll()) <= LML_TOL * abs(o.mll())
    assert np.max(np.abs(gp.factors - np.tril(o.L))) < 1e-10
    assert np.max(np.abs(gp.alpha - o.alpha)) <= 1e-9 * np.max(np.abs(o.alpha))
    g = dsm.grad_mll(gp)
    og = o.grad_mll()
    scale = max(float(o.alpha @ o.alpha), float(n))
    assert np.all(np.abs(g - og) <= GRAD_TOL * np.maximum(np.abs(og), 1e-4 * scale)), (g, og)
    assert np.max(np.abs(gp.alpha - o.alpha)) <= 1e-9 * np.max(np.abs(o.alpha))          # alpha as produced by the gradient path (X^T z)
    xt = np.random.default_rng(n).random((300, D))
    mu, var = gp.prediction(xt)
    omu, ovar = o.prediction(xt)
    assert pred_close(mu, omu, var, ovar, float(np.std(y)))


def test_dsmgp_medium_leaves():
    """144 experts of 300..1300 points: exercises the look-ahead tile schedule across heterogeneous experts."""
    import deepstructuredmixtures_b200 as dsm
    x, y = synth(8000, 8, 21)
    model = dsm.buildDSMGP(x, y, 3, 4, M=150, kernel=dsm.ArdSE(np.zeros(8), 0.0), logNoise=-1.0, rng=21)
    th = np.concatenate([0.2 * np.random.default_rng(2).standard_normal(8), [0.1, -1.0]])
    check_eval(model, th)
    # determinism: the same evaluation twice is bit-identical
    a = model.handle.eval(th)
    b = model.handle.eval(th)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("name", ["cfg1_readme", "cfg2_small", "cfg3_small", "cfg4_small", "cfg5_small"])
def test_golden_vectors_gpu(name):
    """The committed oracle outputs (tests/golden/golden.json) reproduced by the CUDA path."""
    import json, os
    import deepstructuredmixtures_b200 as dsm
    from deepstructuredmixtures_b200 import model as mdl
    from golden.make_golden import CASES, product_kernels, structure, test_points
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))[name]
    case = CASES[name]
    x, y, root, _ = structure(case)
    model = mdl.DSMGP(root, x, y, product_kernels(case), -1.0)
    lml, grad = dsm.evaluate(model, np.array(case["theta"]))
    assert abs(lml - gold["lml"]) <= LML_TOL * abs(gold["lml"])
    og = np.array(gold["grad"])
    assert np.all(np.abs(grad - og) <= GRAD_TOL * np.maximum(np.abs(og), 1e-3 * np.max(np.abs(og)))), (grad, og)
    rows = model.handle.leaf_rows()
    assert relerr(rows[:, 0], np.array(gold["leaf_lml"])) < LML_TOL
    z = dsm.update_(model)
    assert abs(z - gold["z"]) <= LML_TOL * abs(gold["z"])
    mu, var = dsm.predict(model, test_points(case))
    assert relerr(mu, np.array(gold["mu"])) < PRED_TOL and relerr(var, np.array(gold["var"])) < PRED_TOL


def test_streaming_batches_match_resident():
    """keep_factors=0 with a tiny arena: the experts are processed in several batches that reuse the factor arena
    (the cfg5 path: factor -> reduce -> discard).  Must give the same LML and gradient as the resident run."""
    import deepstructuredmixtures_b200 as dsm
    from deepstructuredmixtures_b200 import model as mdl, structure as st
    x, y = synth(6000, 5, 31)
    kern = dsm.ArdSE(np.zeros(5), 0.0)
    cfg = st.DSMGPConfig(None, kern, -1.0, 120, 4, 3, 2, 0.5, True)
    root = st.buildTree(x, y, cfg, np.random.default_rng(31))
    th = np.array([0.1, -0.1, 0.2, 0.0, -0.2, 0.1, -1.0])
    ma = mdl.DSMGP(root, x, y, [kern.copy()], -1.0)
    lml_a, g_a = ma.handle.eval(th)
    ma.close()
    root = st.buildTree(x, y, cfg, np.random.default_rng(31))
    mb = mdl.DSMGP(root, x, y, [kern.copy()], -1.0, keep_factors=False, arena_bytes=24 << 20)
    lml_b, g_b = mb.handle.eval(th)
    assert lml_a == lml_b and np.array_equal(g_a, g_b)
    # the fit-only path (potrf + back-substitution, no inverse) streams the same way and gives the same LML table
    info, _ = mb.handle.fit()
    assert np.all(info == 0)
    assert mb.handle.lml()[mb.handle.tree.root] == lml_a
    with pytest.raises(dsm.DsmgpError):
        dsm.predict(mb, x[:10])
    mb.close()


def test_full_size_cfg3_properties():
    """BASELINE.json config 3 at FULL size (40,000 x 8 ArdSE, V=3 K=4 M=500: 144 experts of 918..5008 points), where the
    oracle would need minutes: size-independent properties instead of an oracle comparison.
      * the factor of a sampled expert reproduces its Gram matrix:  ||L L^T - (K + c I)|| / ||K + c I|| < 1e-13
      * alpha solves the system:                                     ||(K + c I) alpha - y|| / ||y|| < 1e-9
      * mathematical gradients agree with a central finite difference of the LML along a random direction -- up to the
        factor V^2 = 9 of the reference's down-pass, whose sum nodes cancel the -log K prior (optimize.jl:64-74, SURVEY
        App. B Q4): every expert sits below two sum nodes with V = 3 children
      * a second evaluation is bit-identical (deterministic reductions)
      * the root LML is the up-pass of the per-expert LMLs (host restatement of optimize.jl:27-39)
    """
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    import deepstructuredmixtures_b200 as dsm
    from deepstructuredmixtures_b200 import model as mdl
    w = bench.WORKLOADS["cfg3"]
    x, y, root, kern = bench.build_structure(w)
    model = mdl.DSMGP(root, x, y, [kern.copy()], -1.0, as_written_grads=False)
    H = model.handle
    th = bench.thetas([kern.nparams], w["seed"])[1]
    lml, grad, nodes = H.eval(th, want_nodes=True)
    rows = H.leaf_rows()
    assert np.isfinite(lml) and np.all(np.isfinite(grad)) and np.all(H.leaf_info() == 0)
    lml2, grad2 = H.eval(th)
    assert lml2 == lml and np.array_equal(grad, grad2)
    # up-pass of the per-expert LMLs through the oracle's tree walk
    o_root = oracle_tree(model)
    ell = {}
    leaf_lml = {lf.id: rows[lf.leaf_index, 0] for lf in orc.getLeaves(o_root)}

    def up(node):
        if node.type == orc.NODE_LEAF:
            ell[node.id] = leaf_lml[node.id]
        elif node.type == orc.NODE_SPLIT:
            ell[node.id] = sum(up(c) for c in node.children)
        else:
            K = len(node.children)
            ell[node.id] = orc.logsumexp([-math.log(K) + up(c) for c in node.children])
        return ell[node.id]

    assert abs(up(o_root) - lml) <= 1e-12 * abs(lml)
    # finite difference along a random direction (central, step 1e-4: truncation ~1e-8 relative)
    rng = np.random.default_rng(0)
    d = rng.standard_normal(th.size); d /= np.linalg.norm(d)
    hstep = 1e-4
    lp, _ = H.eval(th + hstep * d, want_grad=False)
    lm, _ = H.eval(th - hstep * d, want_grad=False)
    fd = (lp - lm) / (2 * hstep)
    gd = (grad @ d) / w["V"] ** 2
    assert abs(fd - gd) <= 1e-6 * max(abs(fd), np.linalg.norm(grad) / w["V"] ** 2), (fd, gd)
    # factor / alpha of the smallest expert against its Gram matrix (dsmgp_kernelmatrix)
    H.eval(th, want_grad=False)
    sizes = np.diff(H.leaf_ptr)
    l = int(np.argmin(sizes))
    lf = model.leaves[l]
    xl = x[lf.obs - 1]
    k = kern.copy(); k.logl[:] = th[:8]; k.logs = float(th[8])
    F = dsm.kernelmatrix(k, xl, xl)
    # (a corner of the Gram matrix against the formula itself: additive ARD, kernels.jl:31-49)
    sub = xl[:60]
    ell2 = np.exp(th[:8]) ** 2
    Kref = math.exp(2 * th[8]) * np.sum(np.exp(-0.5 * (sub[:, None, :] - sub[None, :, :]) ** 2 / ell2), axis=2)
    assert np.max(np.abs(F[:60, :60] - Kref)) <= 1e-13 * np.max(np.abs(Kref))
    F[np.diag_indices(F.shape[0])] += math.exp(2 * th[9]) + 1e-8
    Lf = H.leaf_factor(l)
    assert np.linalg.norm(Lf @ Lf.T - F) <= 1e-13 * np.linalg.norm(F)
    al = H.leaf_alpha(l)
    yc = y[lf.obs - 1] - lf.mean
    assert np.linalg.norm(F @ al - yc) <= 1e-9 * np.linalg.norm(yc)
    model.close()


def test_full_size_cfg2_properties():
    """BASELINE.json config 2 at full size (10,000 x 1 IsoSE, V=3 K=4 M=100, 144 experts; LAUUM path): finite difference of
    the LML, and the as-written quirk ratio at scale -- kernel gradients carry an extra factor exp(log sigma)
    (kernels.jl:90), the noise gradient does not."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from deepstructuredmixtures_b200 import model as mdl
    w = bench.WORKLOADS["cfg2"]
    x, y, root, kern = bench.build_structure(w)
    th = np.array([-1.2, 0.35, -1.1])
    grads = {}
    for aw in (False, True):
        model = mdl.DSMGP(root, x, y, [kern.copy()], -1.0, as_written_grads=aw)
        lml, grads[aw] = model.handle.eval(th)
        if not aw:
            rng = np.random.default_rng(1)
            d = rng.standard_normal(3); d /= np.linalg.norm(d)
            hstep = 1e-5
            lp, _ = model.handle.eval(th + hstep * d, want_grad=False)
            lm, _ = model.handle.eval(th - hstep * d, want_grad=False)
            fd = (lp - lm) / (2 * hstep)
            gd = (grads[aw] @ d) / w["V"] ** 2
            assert abs(fd - gd) <= 1e-6 * max(abs(fd), np.linalg.norm(grads[aw]) / w["V"] ** 2), (fd, gd)
        model.close()
    s = math.exp(th[1])
    assert np.all(np.abs(grads[True][:2] - s * grads[False][:2]) <= 1e-10 * np.abs(grads[True][:2]))
    assert abs(grads[True][2] - grads[False][2]) <= 1e-12 * abs(grads[False][2])


@pytest.mark.parametrize("opt", ["adam", "rmsprop", "descent"])
def test_train_loop_in_library_matches_host_loop(opt):
    """dsmgp_train (optimisers.jl:40-83 inside the library) == the host loop that calls dsmgp_eval per iteration
    (forced here by passing a callback), for the Flux optimisers incl. the identity-keyed state quirk (App. B Q9)."""
    import deepstructuredmixtures_b200 as dsm
    x, y = synth(600, 2, 41)

    def make():
        return dsm.buildDSMGP(x, y, 2, 2, M=60, kernel=dsm.IsoSE(0.0, 0.0), logNoise=-1.0, rng=41)

    def optimiser(identity):
        return {"adam": dsm.ADAM(0.05, state_by_identity=identity), "rmsprop": dsm.RMSProp(0.05, state_by_identity=identity),
                "descent": dsm.Descent(1e-4, state_by_identity=identity)}[opt]

    for identity in (True, False):
        m1, m2 = make(), make()
        _, ell_lib = dsm.train_(m1, optimiser(identity), iterations=14, randinit=False, lam=1e-12)
        _, ell_host = dsm.train_(m2, optimiser(identity), iterations=14, randinit=False, lam=1e-12, callback=lambda *a: None)
        assert ell_lib.size == ell_host.size == 14
        assert np.max(np.abs(ell_lib - ell_host)) <= 1e-10 * np.max(np.abs(ell_host)), (opt, identity)
        th1 = m1.handle.get_leaf_params(0); th2 = m2.handle.get_leaf_params(0)
        assert np.max(np.abs(th1 - th2)) <= 1e-10
    # early stopping returns the trace up to the stopping iteration in both implementations
    m1, m2 = make(), make()
    _, e1 = dsm.train_(m1, dsm.Descent(1e-9), iterations=40, randinit=False, lam=1e3, earlystop=3)
    _, e2 = dsm.train_(m2, dsm.Descent(1e-9), iterations=40, randinit=False, lam=1e3, earlystop=3, callback=lambda *a: None)
    assert e1.size == e2.size == 13


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_predict_matches_single_handle(world):
    """Leaf-sharded prediction: `world` handles (rank r of world, here all on one GPU) each predict their own experts;
    the SUM of their buffers (what the all-reduce produces) mixed by any rank equals the single-handle prediction."""
    import deepstructuredmixtures_b200 as dsm
    from deepstructuredmixtures_b200 import model as mdl, structure as st
    x, y = synth(3000, 4, 51)
    kern = dsm.ArdSE(np.zeros(4), 0.0)
    cfg = st.DSMGPConfig(None, kern, -1.0, 150, 3, 3, 2, 0.5, True)
    th = np.array([0.1, -0.2, 0.0, 0.2, 0.1, -1.0])
    xt = np.random.default_rng(3).random((900, 4))
    root = st.buildTree(x, y, cfg, np.random.default_rng(51))
    single = mdl.DSMGP(root, x, y, [kern.copy()], -1.0)
    lml0, g0 = single.handle.eval(th); dsm.update_(single)
    mu0, var0 = dsm.predict(single, xt)
    ranks = [mdl.DSMGP(root, x, y, [kern.copy()], -1.0, rank=r, world=world) for r in range(world)]
    # evaluation: every rank fills its rows on the device; their SUM (the all-reduce) goes back into every rank's table
    import torch
    from deepstructuredmixtures_b200.distributed import _DevPtr
    n = single.handle.L * single.handle.row_width
    tens = [torch.as_tensor(_DevPtr(m.handle.eval_local_dev(th), n), device="cuda:0") for m in ranks]
    tot = torch.stack(tens).sum(0)
    for t in tens:
        t.copy_(tot)
    torch.cuda.synchronize()
    bufs = []
    for m in ranks:
        lml, g = m.handle.eval_finish_dev()
        assert lml == lml0 and np.array_equal(g, g0)
        dsm.update_(m)
        bufs.append(m.handle.predict_local(xt))
    total = np.sum(bufs, axis=0)
    nz = [np.count_nonzero(b) for b in bufs]
    assert all(n > 0 for n in nz) and sum(nz) == np.count_nonzero(total)       # disjoint supports
    for m in ranks:
        mu, var = m.handle.predict_finish(xt, total)
        assert np.allclose(mu, mu0, rtol=1e-12, atol=1e-12 * np.max(np.abs(mu0))) and np.allclose(var, var0, rtol=1e-12, atol=0)
        with pytest.raises(dsm.DsmgpError):
            dsm.predict(m, xt)                          # the single-handle call refuses experts of other ranks
    for m in ranks + [single]:
        m.close()
