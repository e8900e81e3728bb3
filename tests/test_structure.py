"""Host-side region-graph construction (treeStructure.jl restated) and overlap matrix (fit.jl:12-39)."""
import numpy as np
import pytest

from conftest import orc, synth


def build(N=3000, D=4, V=3, K=4, M=50, depth=2, eps=0.5, seed=1, kernel=None, useSum=True):
    from deepstructuredmixtures_b200 import kernels as kr, structure as st
    x, y = synth(N, D, seed)
    kern = kr.ArdSE(np.zeros(D), 0.0) if kernel is None else kernel
    cfg = st.DSMGPConfig(None, kern, 1.0, M, K, V, depth, eps, useSum)
    root = st.buildTree(x, y, cfg, np.random.default_rng(seed))
    return x, y, root


def test_leaves_are_ascending_partitions_per_sum_child():
    from deepstructuredmixtures_b200 import structure as st
    x, y, root = build()
    assert isinstance(root, st.GPSumNode) and len(root.children) == 3
    for split in root.children:
        obs = np.concatenate([lf.obs for lf in st.getLeaves(split)])
        # every point appears exactly V(=3) times below a root child (one per nested sum child)
        cnt = np.bincount(obs, minlength=x.shape[0] + 1)[1:]
        assert np.all(cnt == 3)
    for lf in st.getLeaves(root):
        assert np.all(np.diff(lf.obs) > 0) and lf.obs[0] >= 1 and lf.nobs == lf.obs.size
        assert abs(lf.mean - np.mean(y[lf.obs - 1])) < 1e-15
        assert np.all(x[lf.obs - 1] <= lf.ub + 0) and np.all(x[lf.obs - 1] > lf.lb - 0)


def test_leaf_count_and_sizes_match_survey_rules():
    from deepstructuredmixtures_b200 import structure as st
    x, y, root = build(N=4000, M=50, eps=0.0)
    leaves = st.getLeaves(root)
    assert len(leaves) == 144                                  # (V*K)^depth
    sizes = np.array([lf.nobs for lf in leaves])
    assert sizes.sum() == 9 * 4000                             # V^depth * N
    assert sizes.max() - sizes.min() <= 3                      # eps = 0 -> median splits


def test_poe_structure_has_no_sum_nodes():
    from deepstructuredmixtures_b200 import structure as st
    x, y, root = build(N=2000, V=1, K=4, M=100, useSum=False, eps=0.0)

    def rec(n):
        assert not isinstance(n, st.GPSumNode)
        if not isinstance(n, st.GPNode):
            for c in n.children:
                rec(c)
    rec(root)
    obs = np.sort(np.concatenate([lf.obs for lf in st.getLeaves(root)]))
    assert np.array_equal(obs, np.arange(1, 2001))


def test_kernel_mixture_leaves():
    from deepstructuredmixtures_b200 import kernels as kr, structure as st
    x, y, root = build(N=1500, D=3, V=2, K=2, M=100, kernel=[kr.IsoSE(0.0, 0.0), kr.IsoLinear(0.0)])
    ft, leaves = st.flatten(root)
    assert (ft.node_type == 3).sum() * 2 == len(leaves)
    assert [lf.kernelid for lf in leaves[:4]] == [1, 2, 1, 2]


def test_flatten_is_topological_and_routing_thresholds_sorted():
    from deepstructuredmixtures_b200 import structure as st
    x, y, root = build()
    ft, leaves = st.flatten(root)
    for i in range(len(ft.node_type)):
        ch = ft.child_idx[ft.child_ptr[i]:ft.child_ptr[i + 1]]
        assert np.all(ch < i)
        if ft.node_type[i] == 1:
            s = ft.split_val[ft.split_ptr[i]:ft.split_ptr[i + 1]]
            assert len(s) == len(ch) and np.all(np.diff(s) >= 0)
    assert ft.root == len(ft.node_type) - 1
    assert sorted(ft.leaf_of_node[ft.node_type == 0]) == list(range(len(leaves)))


def test_overlap_matrix_matches_oracle_bitset_definition():
    from deepstructuredmixtures_b200 import structure as st
    x, y, root = build(N=1200, D=2, V=2, K=3, M=40)
    ft, leaves = st.flatten(root)
    D = st.getOverlap(root, 1200)
    flat = dict(ft.as_dict())
    flat["leaf_ptr"] = np.concatenate([[0], np.cumsum([lf.nobs for lf in leaves])])
    flat["leaf_obs"] = np.concatenate([lf.obs for lf in leaves])
    flat["leaf_kernel_id"] = np.zeros(len(leaves), dtype=int)
    flat["leaf_mean"] = np.array([lf.mean for lf in leaves])
    oroot = orc.tree_from_flat(flat, x, y, [orc.ArdSE(np.zeros(2), 0.0)], 1.0)
    assert np.array_equal(D, orc.getOverlap(oroot, 1200))
    assert np.all(np.diag(D) == 0) and D.max() <= 1.0 and D.min() >= 0.0
