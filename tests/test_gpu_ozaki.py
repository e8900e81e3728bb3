"""GPU tests of the INT8 split path (csrc/api_ozaki.cu, csrc/ozaki.cuh): experts of >= 8 block rows are split at the middle
block row and the GEMM-shaped products of the factorisation and of the triangular inverse run as error-free INT8 slice products
on the tcgen05 tensor cores.  Every result must agree with the oracle within north_star's tolerances (LML / gradients 1e-9,
predictions 1e-8) AND with the FP64 tile pipelines (DSMGP_OZAKI=0) far more tightly; the tests print the errors they achieve.
The plan is made at dsmgp_create, so the environment switch is set before a handle is built."""
import numpy as np
import pytest

from conftest import orc, relerr, synth

pytestmark = pytest.mark.gpu

LML_TOL = 1e-9
GRAD_TOL = 1e-9
PRED_TOL = 1e-8


def _gp(n, D, seed, kern, monkeypatch, ozaki, as_written=True):
    import deepstructuredmixtures_b200 as dsm
    monkeypatch.setenv("DSMGP_OZAKI", "1" if ozaki else "0")
    x, y = synth(n, D, seed)
    gp = dsm.GaussianProcess(x, y, kernel=kern.copy(), logNoise=-0.7, run_cholesky=True, as_written_grads=as_written)
    return gp, x, y


@pytest.mark.parametrize("n", [1030, 1664, 2500, 4100])
@pytest.mark.parametrize("ktype", ["ardse", "isose"])
def test_single_expert_split_path_against_oracle_and_fp64(ktype, n, monkeypatch):
    """One expert (a 1-leaf handle) through the split path: n = 1030 has a half-wide last block (np = 1088), 1664 = 13 blocks
    (odd split), 2500 / 4100 are cfg3-sized.  LML, as-written and mathematical gradients, alpha and a prediction against the
    oracle, and against the same expert on the FP64 pipelines."""
    import deepstructuredmixtures_b200 as dsm
    D = 3
    kern = dsm.ArdSE([0.1, -0.3, 0.4], 0.2) if ktype == "ardse" else dsm.IsoSE(0.2, 0.1)
    xt = np.random.default_rng(n).random((64, D))
    res = {}
    for oz in (False, True):
        for aw in (True, False):
            gp, x, y = _gp(n, D, 100 + n, kern, monkeypatch, oz, as_written=aw)
            H = dsm.model._model_of(gp).handle
            lml = gp.mll()
            g = dsm.grad_mll(gp)
            info = H.int8_info()
            assert (info["batches"] > 0) == oz, info
            if oz:
                assert info["int8_ops"] > 0 and info["gemm_ms"] > 0
            mu, var = gp.prediction(xt)
            res[(oz, aw)] = (lml, g, np.array(gp.alpha), mu, var)
            dsm.model._model_of(gp).close()
    o = orc.GaussianProcess(x, y, kernel=orc.Kernel(kern.type, kern.logl, kern.logs), logNoise=-0.7, run_cholesky=True)
    scale = max(float(o.alpha @ o.alpha), float(n))
    for aw in (True, False):
        lml, g, alpha, mu, var = res[(True, aw)]
        lml0, g0, alpha0, mu0, var0 = res[(False, aw)]
        og = o.grad_mll() if aw else o.grad_mll(mathematical=True)
        e_lml = abs(lml - o.mll()) / abs(o.mll())
        e_g = float(np.max(np.abs(g - og) / np.maximum(np.abs(og), 1e-4 * scale)))
        d_lml = abs(lml - lml0) / abs(lml0)
        d_g = float(np.max(np.abs(g - g0) / np.maximum(np.abs(g0), 1e-4 * scale)))
        d_a = relerr(alpha, alpha0)
        ea, ea0 = relerr(alpha, o.alpha), relerr(alpha0, o.alpha)     # alpha itself: conditioning-limited in either arithmetic
        omu, ovar = o.prediction(xt)
        e_mu = float(np.max(np.abs(mu - omu) / np.maximum(np.abs(omu), float(np.std(y)))))
        e_var = relerr(var, ovar)
        print(f"\n[int8 split] {ktype} n={n} {'as-written' if aw else 'mathematical'}: vs oracle LML {e_lml:.1e} grad {e_g:.1e} "
              f"predict {e_mu:.1e}/{e_var:.1e}; vs FP64 pipelines LML {d_lml:.1e} grad {d_g:.1e} alpha {d_a:.1e}; "
              f"alpha vs oracle: split {ea:.1e}, FP64 pipelines {ea0:.1e}")
        assert e_lml <= LML_TOL and e_g <= GRAD_TOL and e_mu <= PRED_TOL and e_var <= PRED_TOL
        assert d_lml <= 1e-11 and d_g <= 1e-10
        assert ea <= max(1e-7, 10 * ea0)


def _model(monkeypatch, ozaki, as_written=True, mixture=False):
    import deepstructuredmixtures_b200 as dsm
    monkeypatch.setenv("DSMGP_OZAKI", "1" if ozaki else "0")
    N, D = 9000, 4
    x, y = synth(N, D, 31)
    kern = [dsm.IsoSE(0.1, 0.0), dsm.IsoLinear(0.2)] if mixture else dsm.ArdSE(np.linspace(-0.2, 0.3, D), 0.1)
    return dsm.buildDSMGP(x, y, 2, 2, M=2600, eps=0.3, kernel=kern, rng=5, fit=False, as_written_grads=as_written), x, y


@pytest.mark.parametrize("mixture", [False, True])
def test_model_split_path_equals_fp64_pipelines(mixture, monkeypatch):
    """A 2-level DSMGP whose experts have 1,200 ... 3,000 observations: evaluation (rows, model LML / gradient, per-node LML),
    fit-only path (no inverse: alpha by back-substitution behind the split factorisation), a masked finetune evaluation (the
    factorisation splits, the inverse keeps the masked tile pipeline) and predictions, split path against FP64 pipelines."""
    import deepstructuredmixtures_b200 as dsm
    out = {}
    for oz in (False, True):
        model, x, y = _model(monkeypatch, oz, mixture=mixture)
        H = model.handle
        sizes = np.diff(H.leaf_ptr)
        lml, grad, nodes = H.eval(None, want_nodes=True)
        rows = H.leaf_rows().copy()
        info = H.int8_info()
        assert (info["batches"] > 0) == oz, (info, sizes.max())
        ls = np.zeros(H.L); ls[::2] = 1.0
        lml_m, grad_m = H.eval(None, leaf_scale=ls)
        H.fit()
        lml_fit = H.lml().copy()
        big = int(np.argmax(sizes))
        alpha = H.leaf_alpha(big).copy()
        dsm.update_(model)
        xt = np.random.default_rng(3).random((500, x.shape[1]))
        mu, var = dsm.predict(model, xt)
        out[oz] = dict(lml=lml, grad=grad, nodes=nodes, rows=rows, lml_m=lml_m, grad_m=grad_m, lml_fit=lml_fit, alpha=alpha, mu=mu, var=var,
                       nmax=int(sizes.max()))
        model.close()
    a, b = out[False], out[True]
    gs = np.abs(a["grad"]).max()
    errs = dict(lml=abs(a["lml"] - b["lml"]) / abs(a["lml"]), grad=float(np.abs(a["grad"] - b["grad"]).max() / gs),
                nodes=relerr(b["nodes"], a["nodes"]), rows_lml=relerr(b["rows"][:, 0], a["rows"][:, 0]),
                rows_grad=float(np.abs(a["rows"][:, 1:] - b["rows"][:, 1:]).max() / np.abs(a["rows"][:, 1:]).max()),
                masked=float(np.abs(a["grad_m"] - b["grad_m"]).max() / np.abs(a["grad_m"]).max()),
                fit=relerr(b["lml_fit"], a["lml_fit"]), alpha=relerr(b["alpha"], a["alpha"]),
                mu=float(np.abs(a["mu"] - b["mu"]).max() / np.std(a["mu"])), var=relerr(b["var"], a["var"]))
    print(f"\n[int8 split] model (largest expert {a['nmax']}, mixture={mixture}): " + ", ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    assert errs["lml"] <= 1e-11 and errs["nodes"] <= 1e-11 and errs["rows_lml"] <= 1e-11 and errs["fit"] <= 1e-11
    assert errs["grad"] <= 1e-10 and errs["rows_grad"] <= 1e-10 and errs["masked"] <= 1e-10
    assert errs["alpha"] <= 1e-7 and errs["mu"] <= 1e-9 and errs["var"] <= 1e-9


def test_split_path_switches(monkeypatch):
    """DSMGP_OZAKI_SLICES=7, DSMGP_OZAKI_DEPTH=2, DSMGP_OZAKI_POTRF=0, DSMGP_OZAKI_TRSM=0 and DSMGP_OZAKI_LAUUM=0 are all valid
    configurations of the path: same results within the bounds of their slice count (true gradients: the LAUUM pass runs)."""
    import deepstructuredmixtures_b200 as dsm
    kern = dsm.ArdSE([0.1, -0.3, 0.4], 0.2)
    gp0, x, y = _gp(3300, 3, 77, kern, monkeypatch, False, as_written=False)
    ref = (gp0.mll(), dsm.grad_mll(gp0).copy())
    dsm.model._model_of(gp0).close()
    for name, val, tol in (("DSMGP_OZAKI_SLICES", "7", 1e-8), ("DSMGP_OZAKI_DEPTH", "2", 1e-10), ("DSMGP_OZAKI_POTRF", "0", 1e-10),
                           ("DSMGP_OZAKI_TRSM", "0", 1e-10), ("DSMGP_OZAKI_LAUUM", "0", 1e-10)):
        monkeypatch.setenv(name, val)
        gp, _, _ = _gp(3300, 3, 77, kern, monkeypatch, True, as_written=False)
        lml, g = gp.mll(), dsm.grad_mll(gp)
        assert dsm.model._model_of(gp).handle.int8_info()["batches"] > 0
        e = max(abs(lml - ref[0]) / abs(ref[0]), float(np.abs(g - ref[1]).max() / np.abs(ref[1]).max()))
        print(f"\n[int8 split] {name}={val}: max rel diff to the FP64 pipelines {e:.1e}")
        assert e <= tol
        dsm.model._model_of(gp).close()
        monkeypatch.delenv(name)
