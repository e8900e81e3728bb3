import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# ---- shared helpers (oracle side is imported ONLY here and in the tests) -------------------------
from oracle import dsm_oracle as orc  # noqa: E402


def to_oracle_kernel(k):
    """product KernelFunction -> oracle Kernel"""
    return orc.Kernel(k.type, k.logl.copy(), k.logs)


def oracle_tree(model, theta=None):
    """Rebuild the oracle's node objects from the flat arrays the library consumed."""
    flat = dict(model.flat.as_dict())
    flat["leaf_ptr"] = model.handle.leaf_ptr
    flat["leaf_obs"] = np.concatenate([lf.obs for lf in model.leaves])
    flat["leaf_kernel_id"] = np.array([lf.kernelid - 1 for lf in model.leaves])
    flat["leaf_mean"] = np.array([lf.mean for lf in model.leaves])
    kernels = [to_oracle_kernel(k) for k in model.kernels]
    root = orc.tree_from_flat(flat, model.x, model.y, kernels, model.leaves[0].logNoise)
    if theta is not None:
        orc.setparams(root, theta)
    return root


def synth(N, D, seed, sorted1d=False):
    """SURVEY §8(d) synthetic inputs: x ~ U[0,1)^{N x D}, y = sin(2 pi x.w) + 0.1 eps."""
    rng = np.random.default_rng(seed)
    x = rng.random((N, D))
    if sorted1d and D == 1:
        x = np.sort(x, axis=0)
    w = rng.standard_normal(D)
    y = np.sin(2 * np.pi * (x @ w)) + 0.1 * rng.standard_normal(N)
    return x, y


def theta0(kernels, logl=0.0, logs=0.0, logn=-1.0):
    out = []
    for k in kernels:
        out.extend([logl] * k.logl.size + [logs, logn])
    return np.array(out, dtype=np.float64)


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0
