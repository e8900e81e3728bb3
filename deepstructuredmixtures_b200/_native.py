"""ctypes binding of libdsmgp.so (include/dsmgp.h).

This is the Python twin of the Julia `ccall` shim in julia/DSMGPNative.jl: one thin wrapper per
C entry point, pointer passing only.  The library is REQUIRED: there is no Python/NumPy fallback,
and a missing or unloadable `libdsmgp.so` raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DSMGP_LIB_PATH: development override for A/B runs of kernel variants (still a CUDA build of this library)
LIB_PATH = os.environ.get("DSMGP_LIB_PATH") or os.path.join(_HERE, "libdsmgp.so")

OK, ERR_ARG, ERR_CUDA, ERR_COMM, ERR_OOM, ERR_NOT_PD, ERR_STATE = range(7)
ISO_SE, ARD_SE, ISO_LINEAR, ARD_LINEAR = 0, 1, 2, 3
NODE_LEAF, NODE_SPLIT, NODE_SUM, NODE_KSUM = 0, 1, 2, 3
PREDICT_DSMGP, PREDICT_POE, PREDICT_GPOE, PREDICT_RBCM = 0, 1, 2, 3


class DsmgpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libdsmgp error {code}: {msg}")
        self.code = code


class PosDefException(DsmgpError):
    """Strict mode: a leaf's Gram matrix was not positive definite (LinearAlgebra.PosDefException)."""


class KernelDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("nparams", C.c_int32)]


class Tree(C.Structure):
    _fields_ = [("n_nodes", C.c_int64), ("node_type", C.POINTER(C.c_int32)),
                ("child_ptr", C.POINTER(C.c_int64)), ("child_idx", C.POINTER(C.c_int64)),
                ("leaf_of_node", C.POINTER(C.c_int64)), ("split_dim", C.POINTER(C.c_int32)),
                ("split_ptr", C.POINTER(C.c_int64)), ("split_val", C.POINTER(C.c_double)),
                ("root", C.c_int64)]


class Opts(C.Structure):
    _fields_ = [("as_written_grads", C.c_int32), ("keep_factors", C.c_int32), ("rank", C.c_int32),
                ("world", C.c_int32), ("device", C.c_int32), ("strict_pd", C.c_int32),
                ("arena_bytes", C.c_int64), ("reserved", C.c_int32 * 8)]


class Timings(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("gram_ms", "potrf_ms", "solve_ms", "inverse_ms", "grad_ms", "tree_ms",
                                          "total_ms", "potrf_flops", "inverse_flops", "gram_bytes")] + [("launches", C.c_int64)] + \
               [(n, C.c_double) for n in ("predict_ms", "predict_flops", "predict_bytes")]


# every symbol include/dsmgp.h declares (tests/test_abi.py checks the list against the header)
EXPORTS = [
    "dsmgp_default_opts", "dsmgp_create", "dsmgp_destroy", "dsmgp_last_error", "dsmgp_set_params",
    "dsmgp_set_leaf_params", "dsmgp_get_leaf_params", "dsmgp_nparams", "dsmgp_n_leaves", "dsmgp_n_nodes",
    "dsmgp_leaf_size", "dsmgp_fit", "dsmgp_lml", "dsmgp_grad", "dsmgp_eval", "dsmgp_row_width",
    "dsmgp_finetune_eval", "dsmgp_train", "dsmgp_eval_local_dev", "dsmgp_eval_finish_dev", "dsmgp_leaf_rows", "dsmgp_leaf_owner",
    "dsmgp_update_weights", "dsmgp_predict", "dsmgp_predict_local", "dsmgp_predict_finish", "dsmgp_leaf_predict", "dsmgp_leaf_alpha", "dsmgp_leaf_factor",
    "dsmgp_leaf_info", "dsmgp_kernelmatrix", "dsmgp_overlap", "dsmgp_release_cache", "dsmgp_chol_continue", "dsmgp_chol_delete_rows", "dsmgp_potrf",
    "dsmgp_host_tree_eval", "dsmgp_host_shard", "dsmgp_get_timings", "dsmgp_set_profiling", "dsmgp_int8_info", "dsmgp_host_split_plan",
    "dsmgp_set_sharing", "dsmgp_get_sharing", "dsmgp_infer", "dsmgp_reset_weights", "dsmgp_comm_unique_id", "dsmgp_comm_init", "dsmgp_host_sharing_plan",
    "dsmgp_part_create", "dsmgp_part_destroy", "dsmgp_part_size", "dsmgp_part_range", "dsmgp_part_sorted_column", "dsmgp_part_split",
    "dsmgp_part_rows", "dsmgp_overlap_csr", "dsmgp_chol_delete_rows_batched",
]
COMM_ID_BYTES = 128

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load libdsmgp.so (built in-tree by `__graft_entry__.build()` / `make -C deepstructuredmixtures_b200/csrc`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                          "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    P, I32, I64, D = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    pd, pi32, pi64 = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    sig = {
        "dsmgp_default_opts": (None, [C.POINTER(Opts)]),
        "dsmgp_create": (I32, [pd, I64, I64, I64, pi64, pi64, pd, pd, pi32, C.POINTER(KernelDesc), I32,
                               C.POINTER(Tree), C.POINTER(Opts), C.POINTER(P)]),
        "dsmgp_destroy": (None, [P]),
        "dsmgp_last_error": (C.c_char_p, [P]),
        "dsmgp_set_params": (I32, [P, pd, I64]),
        "dsmgp_set_leaf_params": (I32, [P, I64, pd, I64]),
        "dsmgp_get_leaf_params": (I32, [P, I64, pd, I64]),
        "dsmgp_nparams": (I64, [P]), "dsmgp_n_leaves": (I64, [P]), "dsmgp_n_nodes": (I64, [P]),
        "dsmgp_leaf_size": (I64, [P, I64]),
        "dsmgp_fit": (I32, [P, D, pd, pi32, pd]),
        "dsmgp_set_sharing": (I32, [P, pd, D]),
        "dsmgp_get_sharing": (I32, [P, pi32, pi32, pi32]),
        "dsmgp_infer": (I32, [P, pd, pd]),
        "dsmgp_reset_weights": (I32, [P, pd]),
        "dsmgp_comm_unique_id": (I32, [C.c_void_p]),
        "dsmgp_comm_init": (I32, [P, C.c_void_p]),
        "dsmgp_lml": (I32, [P, pd]),
        "dsmgp_grad": (I32, [P, pd, pd]),
        "dsmgp_eval": (I32, [P, pd, I64, pd, pd, pd, pd]),
        "dsmgp_finetune_eval": (I32, [P, I64, pi64, pd, pd, pd, pd, pd]),
        "dsmgp_predict_local": (I32, [P, pd, I64, I32, pd, pi64]),
        "dsmgp_predict_finish": (I32, [P, pd, I64, I32, pd, pd, pd]),
        "dsmgp_train": (I32, [P, I32, D, D, D, I32, I64, D, I64, pd, pd, pi64]),
        "dsmgp_overlap": (I32, [I64, I64, pi64, pi64, pi32, C.POINTER(Tree), pd]),
        "dsmgp_overlap_csr": (I32, [I64, I64, pi64, pi64, pi32, C.POINTER(Tree), pi64, pi32, pd]),
        "dsmgp_release_cache": (None, []),
        "dsmgp_row_width": (I64, [P]),
        "dsmgp_eval_local_dev": (I32, [P, pd, I64, C.POINTER(C.c_void_p)]),
        "dsmgp_eval_finish_dev": (I32, [P, pd, pd, pd, pd]),
        "dsmgp_leaf_rows": (I32, [P, pd]),
        "dsmgp_leaf_owner": (I32, [P, pi32]),
        "dsmgp_update_weights": (I32, [P, pd, pd]),
        "dsmgp_predict": (I32, [P, pd, I64, I32, pd, pd]),
        "dsmgp_leaf_predict": (I32, [P, I64, pd, I64, pd, pd]),
        "dsmgp_leaf_alpha": (I32, [P, I64, pd]),
        "dsmgp_leaf_factor": (I32, [P, I64, pd]),
        "dsmgp_leaf_info": (I32, [P, pi32]),
        "dsmgp_kernelmatrix": (I32, [I32, pd, I64, pd, I64, pd, I64, pd]),
        "dsmgp_chol_continue": (I32, [pd, I64, I64, pi32]),
        "dsmgp_chol_delete_rows": (I32, [pd, I64, pi64, I64, pd]),
        "dsmgp_chol_delete_rows_batched": (I32, [I64, C.POINTER(pd), pi64, C.POINTER(pi64), pi64, C.POINTER(pd)]),
        "dsmgp_potrf": (I32, [pd, I64, pi32]),
        "dsmgp_host_tree_eval": (I32, [C.POINTER(Tree), I64, pi32, C.POINTER(KernelDesc), I32, pd, I64, pd, pd, pd, pd, pd]),
        "dsmgp_host_shard": (I32, [I64, pi64, I32, pi32]),
        "dsmgp_part_create": (I32, [pd, I64, I64, C.POINTER(P)]),
        "dsmgp_part_destroy": (None, [P]),
        "dsmgp_part_size": (I64, [P, I64]),
        "dsmgp_part_range": (I32, [P, I64, pd, pd]),
        "dsmgp_part_sorted_column": (I32, [P, I64, I64, pd]),
        "dsmgp_part_split": (I32, [P, I64, I64, pd, pd, I64, pi64, pi64]),
        "dsmgp_part_rows": (I32, [P, I64, pi64]),
        "dsmgp_host_sharing_plan": (I32, [I64, pi64, pi64, pi32, pd, D, pi32, pi32, pi32]),
        "dsmgp_get_timings": (I32, [P, C.POINTER(Timings)]),
        "dsmgp_set_profiling": (I32, [P, I32]),
        "dsmgp_int8_info": (I32, [P, pd, I32]),
        "dsmgp_host_split_plan": (I32, [I64, I32, I32, pi32, pi32, pd]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


def f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def colmajor(a) -> np.ndarray:
    """2-D array -> Fortran-ordered float64 (Julia layout)."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def p_d(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def p_i32(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def p_i64(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int64))


def check(rc: int, handle=None):
    if rc == OK:
        return
    msg = lib().dsmgp_last_error(handle)
    msg = msg.decode() if msg else ""
    if rc == ERR_NOT_PD:
        raise PosDefException(rc, msg)
    raise DsmgpError(rc, msg)


class FlatTree:
    """Owns the NumPy arrays behind a `dsmgp_tree` struct."""

    def __init__(self, node_type, child_ptr, child_idx, leaf_of_node, split_dim, split_ptr, split_val, root):
        self.node_type = np.ascontiguousarray(node_type, dtype=np.int32)
        self.child_ptr = np.ascontiguousarray(child_ptr, dtype=np.int64)
        self.child_idx = np.ascontiguousarray(child_idx, dtype=np.int64)
        self.leaf_of_node = np.ascontiguousarray(leaf_of_node, dtype=np.int64)
        self.split_dim = np.ascontiguousarray(split_dim, dtype=np.int32)
        self.split_ptr = np.ascontiguousarray(split_ptr, dtype=np.int64)
        self.split_val = np.ascontiguousarray(split_val, dtype=np.float64)
        if self.split_val.size == 0:
            self.split_val = np.zeros(1)
        if self.child_idx.size == 0:
            self.child_idx = np.zeros(1, dtype=np.int64)
        self.root = int(root)
        self.struct = Tree(len(self.node_type), p_i32(self.node_type), p_i64(self.child_ptr), p_i64(self.child_idx),
                           p_i64(self.leaf_of_node), p_i32(self.split_dim), p_i64(self.split_ptr), p_d(self.split_val),
                           self.root)

    def as_dict(self) -> dict:
        return dict(node_type=self.node_type, child_ptr=self.child_ptr, child_idx=self.child_idx,
                    leaf_of_node=self.leaf_of_node, split_dim=self.split_dim, split_ptr=self.split_ptr,
                    split_val=self.split_val, root=self.root)
