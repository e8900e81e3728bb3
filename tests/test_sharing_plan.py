"""The sharing plan of fit! (fit.jl:71-122): the library's host-side plan (dsmgp_host_sharing_plan, what dsmgp_fit(tau, overlap)
stores) against the oracle's restatement of the reference's scheduling and case split (oracle.fit_plan).  CPU only."""
import numpy as np
import pytest

from conftest import orc

BLK = 128


def _build(N, D, V, K, M, eps, seed, sorted1d=False, mixture=False):
    from deepstructuredmixtures_b200 import kernels as kr, structure as st
    from conftest import synth
    x, y = synth(N, D, seed, sorted1d=sorted1d)
    kern = [kr.IsoSE(0.0, 0.0), kr.IsoLinear(0.0)] if mixture else kr.IsoSE(0.0, 0.0)
    cfg = st.DSMGPConfig(None, kern, -1.0, M, K, V, 2, eps, True)
    root = st.buildTree(x, y, cfg, np.random.default_rng(seed))
    ft, leaves = st.flatten(root)
    Dm = st.getOverlap(root, N)
    return x, y, root, ft, leaves, Dm, (kern if isinstance(kern, list) else [kern])


def _library_plan(leaves, Dm, tau):
    import ctypes as C
    from deepstructuredmixtures_b200 import _native as nat
    L = len(leaves)
    lp = np.zeros(L + 1, dtype=np.int64); lp[1:] = np.cumsum([lf.nobs for lf in leaves])
    obs = np.ascontiguousarray(np.concatenate([lf.obs for lf in leaves]), dtype=np.int64)
    kid = np.ascontiguousarray([lf.kernelid - 1 for lf in leaves], dtype=np.int32)
    ov = np.asfortranarray(Dm, dtype=np.float64)
    kind = np.zeros(L, dtype=np.int32); src = np.zeros(L, dtype=np.int32); blocks = np.zeros(L, dtype=np.int32)
    nat.check(nat.lib().dsmgp_host_sharing_plan(L, nat.p_i64(lp), nat.p_i64(obs), nat.p_i32(kid), nat.p_d(ov), float(tau),
                                                nat.p_i32(kind), nat.p_i32(src), nat.p_i32(blocks)))
    return kind, src, blocks


def _oracle_root(x, y, ft, leaves, kernels):
    flat = dict(ft.as_dict())
    flat["leaf_ptr"] = np.concatenate([[0], np.cumsum([lf.nobs for lf in leaves])])
    flat["leaf_obs"] = np.concatenate([lf.obs for lf in leaves])
    flat["leaf_kernel_id"] = np.array([lf.kernelid - 1 for lf in leaves])
    flat["leaf_mean"] = np.array([lf.mean for lf in leaves])
    return orc.tree_from_flat(flat, x, y, [orc.Kernel(k.type, k.logl, k.logs) for k in kernels], -1.0)


@pytest.mark.parametrize("case", ["sorted1d", "eps0", "scattered", "mixture"])
@pytest.mark.parametrize("tau", [0.05, 0.5])
def test_plan_follows_the_reference_case_split(case, tau):
    cfgs = {"sorted1d": dict(N=6000, D=1, V=3, K=4, M=100, eps=0.5, seed=2, sorted1d=True),
            "eps0": dict(N=4000, D=1, V=3, K=3, M=200, eps=0.0, seed=5, sorted1d=True),
            "scattered": dict(N=3000, D=4, V=3, K=3, M=100, eps=0.5, seed=7),
            "mixture": dict(N=3000, D=1, V=2, K=3, M=150, eps=0.0, seed=9, sorted1d=True, mixture=True)}
    x, y, root, ft, leaves, Dm, kernels = _build(**cfgs[case])
    kind, src, blocks = _library_plan(leaves, Dm, tau)
    oplan = orc.fit_plan(_oracle_root(x, y, ft, leaves, kernels), Dm, tau)
    n_alias = n_prefix = 0
    for j, (branch, main) in enumerate(oplan):
        oj = leaves[j].obs
        if branch == "copy":
            assert kind[j] == 1, (j, branch)
            assert np.array_equal(leaves[src[j]].obs, oj) and leaves[src[j]].kernelid == leaves[j].kernelid
            assert kind[src[j]] != 1                      # sources are resolved through aliases
            n_alias += 1
        elif branch in ("delete", "continue"):
            s = main if kind[main] != 1 else src[main]
            os_ = leaves[s].obs
            k = 0
            while k < min(len(oj), len(os_)) and oj[k] == os_[k]:
                k += 1
            if k // BLK >= 1 and kind[s] != 2:
                assert kind[j] == 2 and src[j] == s and blocks[j] == k // BLK, (j, branch, k)
                n_prefix += 1
            else:
                assert kind[j] == 0
        else:
            assert kind[j] == 0, (j, branch, kind[j])
    if case == "eps0":
        assert n_alias > 0           # identical experts under different sum children (same split dimension, eps = 0)
    if case == "sorted1d" and tau == 0.5:
        assert n_prefix > 0          # contiguous 1-D regions share their leading observations
    if case == "scattered":
        assert n_alias == 0 and n_prefix == 0
    if case == "mixture":            # D[n,m] == 1 for experts of different kernels (fit.jl:28-31) must never alias them
        for j in range(len(leaves)):
            if kind[j] != 0:
                assert leaves[src[j]].kernelid == leaves[j].kernelid


def test_plan_is_empty_without_overlap_information():
    x, y, root, ft, leaves, Dm, kernels = _build(N=2000, D=1, V=2, K=3, M=150, eps=0.0, seed=3, sorted1d=True)
    kind, src, blocks = _library_plan(leaves, np.zeros_like(Dm), 0.05)        # D == 0 (PoE models, fit.jl:78: main = leaf 1)
    assert not kind.any() and (src == -1).all() and not blocks.any()
