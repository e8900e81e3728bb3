"""Generates tests/golden/golden.json: oracle outputs on small seeded cases of the five BASELINE configs' shapes.
Run `python tests/golden/make_golden.py` to regenerate (only when the oracle changes on purpose).
The reference itself cannot be executed here (no Julia), so these vectors pin the ORACLE, and through it the GPU path."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from oracle import dsm_oracle as orc  # noqa: E402

# name: (N, D, kernels, V, K, M, depth, eps, seed, theta)
CASES = {
    "cfg1_readme": dict(N=100, D=1, kernels=["isose"], V=3, K=4, M=10, depth=2, eps=0.5, seed=1, theta=[-1.0, 0.2, -1.2]),
    "cfg2_small": dict(N=600, D=1, kernels=["isose"], V=3, K=4, M=20, depth=2, eps=0.5, seed=2, theta=[-2.0, 0.1, -1.0]),
    "cfg3_small": dict(N=900, D=8, kernels=["ardse"], V=3, K=4, M=30, depth=2, eps=0.5, seed=3,
                       theta=[0.1, -0.1, 0.2, 0.0, 0.3, -0.2, 0.1, 0.0, 0.1, -1.0]),
    "cfg4_small": dict(N=800, D=9, kernels=["isose", "isolin"], V=2, K=4, M=60, depth=2, eps=0.5, seed=4,
                       theta=[0.5, 0.1, -1.0, 1.0, 0.0, -0.8]),
    "cfg5_small": dict(N=1500, D=8, kernels=["ardse"], V=2, K=2, M=40, depth=3, eps=0.1, seed=5,
                       theta=[0.2, 0.1, 0.0, -0.1, 0.3, 0.2, 0.1, 0.0, -0.1, -1.1]),
}


def product_kernels(case):
    from deepstructuredmixtures_b200 import kernels as kr
    out = []
    for k in case["kernels"]:
        out.append({"isose": lambda: kr.IsoSE(0.0, 0.0), "ardse": lambda: kr.ArdSE(np.zeros(case["D"]), 0.0),
                    "isolin": lambda: kr.IsoLinear(0.0)}[k]())
    return out


def oracle_kernels(case):
    return [{"isose": lambda: orc.IsoSE(0.0, 0.0), "ardse": lambda: orc.ArdSE(np.zeros(case["D"]), 0.0),
             "isolin": lambda: orc.IsoLinear(0.0)}[k]() for k in case["kernels"]]


def data(case):
    rng = np.random.default_rng(case["seed"])
    if case["N"] == 100 and case["D"] == 1:      # README example (README.md:33-38)
        x = np.linspace(0, 1, 100).reshape(-1, 1)
        y = np.sin(x[:, 0] * 4 * np.pi + rng.standard_normal(100) * 0.2)
        return x, y
    x = rng.random((case["N"], case["D"]))
    if case["D"] == 1:
        x = np.sort(x, axis=0)
    w = rng.standard_normal(case["D"])
    y = np.sin(2 * np.pi * (x @ w)) + 0.1 * rng.standard_normal(case["N"])
    return x, y


def structure(case):
    """host tree (product builder; partitions are inputs of both the library and the oracle)"""
    from deepstructuredmixtures_b200 import structure as st
    x, y = data(case)
    pk = product_kernels(case)
    cfg = st.DSMGPConfig(None, pk if len(pk) > 1 else pk[0], -1.0, case["M"], case["K"], case["V"], case["depth"],
                         case["eps"], True)
    root = st.buildTree(x, y, cfg, np.random.default_rng(case["seed"]))
    ft, leaves = st.flatten(root)
    flat = dict(ft.as_dict())
    flat["leaf_ptr"] = np.concatenate([[0], np.cumsum([lf.nobs for lf in leaves])])
    flat["leaf_obs"] = np.concatenate([lf.obs for lf in leaves])
    flat["leaf_kernel_id"] = np.array([lf.kernelid - 1 for lf in leaves])
    flat["leaf_mean"] = np.array([lf.mean for lf in leaves])
    return x, y, root, flat


def test_points(case):
    return np.random.default_rng(100 + case["seed"]).random((64, case["D"]))


def run_case(name):
    case = CASES[name]
    x, y, _, flat = structure(case)
    root = orc.tree_from_flat(flat, x, y, oracle_kernels(case), -1.0)
    lml, grad, ell, rows = orc.evaluate(root, case["theta"])
    z = orc.update_weights(root)
    mu, var = orc.predict_dsmgp(root, test_points(case))
    return {"lml": float(lml), "grad": [float(g) for g in grad], "z": float(z), "mu": mu.tolist(), "var": var.tolist(),
            "leaf_lml": [float(rows[l][0]) for l in sorted(rows)], "n_leaves": len(rows)}


if __name__ == "__main__":
    out = {name: run_case(name) for name in CASES}
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f)
    print({k: (v["n_leaves"], v["lml"]) for k, v in out.items()})
