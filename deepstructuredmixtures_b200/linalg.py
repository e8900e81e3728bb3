"""Stand-alone dense operators of the `AdvancedCholesky` sub-module (src/AdvancedCholeskey.jl), on the GPU."""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import numpy as np

from . import _native as nat


def potrf_(A: np.ndarray) -> Tuple[np.ndarray, int]:
    """LAPACK.potrf!('L', A) as used at gaussianprocess.jl:101: returns (lower factor, info)."""
    A = np.array(A, dtype=np.float64, order="F")
    info = C.c_int32(0)
    nat.check(nat.lib().dsmgp_potrf(nat.p_d(A), A.shape[0], C.byref(info)))
    return A, info.value


def chol_continue_(A: np.ndarray, ki: int) -> Tuple[np.ndarray, int]:
    """chol_continue!(A, ki) AdvancedCholeskey.jl:152-174 (ki 1-based): rows/cols < ki hold a valid lower factor."""
    A = np.array(A, dtype=np.float64, order="F")
    info = C.c_int32(0)
    nat.check(nat.lib().dsmgp_chol_continue(nat.p_d(A), A.shape[0], int(ki), C.byref(info)))
    return A, info.value


def chol_delete_rows(Lf: np.ndarray, rows: Sequence[int]) -> np.ndarray:
    """Factor of A[keep, keep] from the factor of A (the row-deletion fit.jl:179-195 builds from lowrankupdate!,
    AdvancedCholeskey.jl:20-59 -- implemented correctly).  `rows` 1-based ascending."""
    Lf = np.array(Lf, dtype=np.float64, order="F")
    rows = np.ascontiguousarray(sorted(rows), dtype=np.int64)
    n = Lf.shape[0]
    out = np.zeros((n - rows.size, n - rows.size), order="F")
    nat.check(nat.lib().dsmgp_chol_delete_rows(nat.p_d(Lf), n, nat.p_i64(rows), rows.size, nat.p_d(out)))
    return out


def overlap_matrix(N: int, leaf_obs: Sequence[np.ndarray], leaf_kernel_id: Sequence[int], tree: "nat.FlatTree") -> np.ndarray:
    """getOverlap(spn, D, gpmap) fit.jl:12-39 on the device (`dsmgp_overlap`): the L x L overlap matrix D."""
    L = len(leaf_obs)
    leaf_ptr = np.zeros(L + 1, dtype=np.int64)
    leaf_ptr[1:] = np.cumsum([len(o) for o in leaf_obs])
    obs = np.ascontiguousarray(np.concatenate(leaf_obs), dtype=np.int64)
    kid = np.ascontiguousarray(leaf_kernel_id, dtype=np.int32)
    D = np.zeros((L, L), order="F")
    nat.check(nat.lib().dsmgp_overlap(int(N), L, nat.p_i64(leaf_ptr), nat.p_i64(obs), nat.p_i32(kid),
                                      C.byref(tree.struct), nat.p_d(D)))
    return D


def overlap_matrix_csr(N: int, leaf_obs: Sequence[np.ndarray], leaf_kernel_id: Sequence[int], tree: "nat.FlatTree"):
    """getOverlap in CSR form (`dsmgp_overlap_csr`): (row_ptr[L+1], col[nnz], val[nnz]) -- only the non-zero entries leave the
    device, for models whose dense L x L matrix is too large to move."""
    L = len(leaf_obs)
    leaf_ptr = np.zeros(L + 1, dtype=np.int64)
    leaf_ptr[1:] = np.cumsum([len(o) for o in leaf_obs])
    obs = np.ascontiguousarray(np.concatenate(leaf_obs), dtype=np.int64)
    kid = np.ascontiguousarray(leaf_kernel_id, dtype=np.int32)
    row_ptr = np.zeros(L + 1, dtype=np.int64)
    args = (int(N), L, nat.p_i64(leaf_ptr), nat.p_i64(obs), nat.p_i32(kid), C.byref(tree.struct), nat.p_i64(row_ptr))
    nat.check(nat.lib().dsmgp_overlap_csr(*args, None, None))
    nnz = int(row_ptr[-1])
    col = np.zeros(max(nnz, 1), dtype=np.int32); val = np.zeros(max(nnz, 1))
    if nnz:
        nat.check(nat.lib().dsmgp_overlap_csr(*args, nat.p_i32(col), nat.p_d(val)))
    return row_ptr, col[:nnz], val[:nnz]


def chol_delete_rows_batched(factors: Sequence[np.ndarray], rows: Sequence[Sequence[int]]):
    """`dsmgp_chol_delete_rows_batched`: row deletion for several factors in one launch (one CTA per matrix, one column sweep per
    matrix for all of its deleted rows).  Returns the list of reduced factors."""
    Ls = [np.array(L, dtype=np.float64, order="F") for L in factors]
    rs = [np.ascontiguousarray(sorted(r), dtype=np.int64) for r in rows]
    cnt = len(Ls)
    n = np.array([L.shape[0] for L in Ls], dtype=np.int64)
    nr = np.array([r.size for r in rs], dtype=np.int64)
    outs = [np.zeros((int(n[m] - nr[m]), int(n[m] - nr[m])), order="F") for m in range(cnt)]
    PD, PI = C.POINTER(C.c_double), C.POINTER(C.c_int64)
    a_arr = (PD * cnt)(*[nat.p_d(L) for L in Ls])
    r_arr = (PI * cnt)(*[nat.p_i64(r) if r.size else C.cast(None, PI) for r in rs])
    o_arr = (PD * cnt)(*[nat.p_d(o) for o in outs])
    nat.check(nat.lib().dsmgp_chol_delete_rows_batched(cnt, a_arr, nat.p_i64(n), r_arr, nat.p_i64(nr), o_arr))
    return outs
