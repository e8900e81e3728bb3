"""The C-ABI library loads, exports every symbol include/dsmgp.h declares, refuses to compute without a GPU
(no CPU fallback), and its HOST-ONLY helpers (tree passes, sharding) agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu, orc


def header_symbols():
    src = open(os.path.join(ROOT, "include", "dsmgp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dsmgp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from deepstructuredmixtures_b200 import _native as nat
    lib = nat.lib()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dsmgp.h but not exported"
    assert sorted(nat.EXPORTS) == syms


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "deepstructuredmixtures_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("dsm_oracle", "oracle") or "import" not in "".join(
                    l for l in txt.splitlines() if "oracle" in l), f
    for l in open(os.path.join(ROOT, "deepstructuredmixtures_b200", "__init__.py")):
        assert "oracle" not in l


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    import deepstructuredmixtures_b200 as dsm
    with pytest.raises(dsm.DsmgpError) as e:
        dsm.kernelmatrix(dsm.IsoSE(0.0, 0.0), np.zeros((4, 1)))
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)
    with pytest.raises(dsm.DsmgpError):
        dsm.GaussianProcess(np.random.rand(10, 1), np.random.rand(10))
    with pytest.raises(dsm.DsmgpError):
        dsm.potrf_(np.eye(4))


def test_create_validates_arguments():
    from deepstructuredmixtures_b200 import _native as nat
    lib = nat.lib()
    h = C.c_void_p()
    rc = lib.dsmgp_create(None, 10, 1, 1, None, None, None, None, None, None, 1, None, None, C.byref(h))
    assert rc == nat.ERR_ARG and b"null" in lib.dsmgp_last_error(None)


def _structure(seed=3, mixture=False):
    from deepstructuredmixtures_b200 import kernels as kr, structure as st
    from conftest import synth
    x, y = synth(700, 3, seed)
    kern = [kr.IsoSE(0.0, 0.0), kr.IsoLinear(0.0)] if mixture else kr.ArdSE(np.zeros(3), 0.0)
    cfg = st.DSMGPConfig(None, kern, -1.0, 40, 3, 2, 2, 0.5, True)
    root = st.buildTree(x, y, cfg, np.random.default_rng(seed))
    ft, leaves = st.flatten(root)
    return x, y, root, ft, leaves, (kern if isinstance(kern, list) else [kern])


def _oracle_root(x, y, ft, leaves, kernels):
    flat = dict(ft.as_dict())
    flat["leaf_ptr"] = np.concatenate([[0], np.cumsum([lf.nobs for lf in leaves])])
    flat["leaf_obs"] = np.concatenate([lf.obs for lf in leaves])
    flat["leaf_kernel_id"] = np.array([lf.kernelid - 1 for lf in leaves])
    flat["leaf_mean"] = np.array([lf.mean for lf in leaves])
    return orc.tree_from_flat(flat, x, y, [orc.Kernel(k.type, k.logl, k.logs) for k in kernels], -1.0)


@pytest.mark.parametrize("mixture", [False, True])
def test_host_tree_passes_match_oracle(mixture):
    """dsmgp_host_tree_eval (optimize.jl:27-150, common.jl:323-334) on oracle-computed leaf rows."""
    from deepstructuredmixtures_b200 import distributed as dd
    x, y, root, ft, leaves, kernels = _structure(4, mixture)
    oroot = _oracle_root(x, y, ft, leaves, kernels)
    theta = np.array([0.2, 0.1, -1.0, 0.3, 0.0, -0.8]) if mixture else np.array([0.1, -0.2, 0.3, 0.1, -1.0])
    lml, grad, ell, rows = orc.evaluate(oroot, theta)
    z = orc.update_weights(oroot)
    Hmax = max(k.nparams for k in kernels)
    tab = np.zeros((len(leaves), 1 + Hmax))
    for l, r in rows.items():
        tab[l, :r.size] = r
    node_lml, g, lw, zz = dd.host_tree_eval(ft, [lf.kernelid - 1 for lf in leaves], kernels, tab)
    assert abs(node_lml[ft.root] - lml) <= 1e-13 * abs(lml)
    for nid, v in ell.items():
        assert abs(node_lml[nid] - v) <= 1e-13 * max(1.0, abs(v))
    assert np.allclose(g, grad, rtol=1e-12, atol=1e-12)
    assert abs(zz - z) <= 1e-13 * abs(z)
    # finetune variant: leaf weights D[g,:]
    D = orc.getOverlap(oroot, x.shape[0])
    _, grad2, _, _ = orc.evaluate(oroot, theta, Drow=D[2])
    _, g2, _, _ = dd.host_tree_eval(ft, [lf.kernelid - 1 for lf in leaves], kernels, tab, leaf_scale=D[2])
    assert np.allclose(g2, grad2, rtol=1e-12, atol=1e-12)


def test_host_shard_is_balanced_and_deterministic():
    from deepstructuredmixtures_b200 import distributed as dd
    rng = np.random.default_rng(0)
    n = rng.integers(200, 6000, size=144)
    lp = np.concatenate([[0], np.cumsum(n)])
    for world in (1, 2, 4, 8):
        o1 = dd.shard_leaves(lp, world)
        o2 = dd.shard_leaves(lp, world)
        assert np.array_equal(o1, o2) and o1.min() == 0 and o1.max() == world - 1
        load = np.array([np.sum(n[o1 == r].astype(float) ** 3) for r in range(world)])
        assert load.max() / load.mean() < 1.15
