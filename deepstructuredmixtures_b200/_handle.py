"""Thin object wrapper around a `dsmgp_handle*` (one per model / per GaussianProcess)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _native as nat
from .kernels import KernelFunction


def comm_unique_id() -> bytes:
    """dsmgp_comm_unique_id: NCCL's 128-byte id, created by rank 0 and handed to every rank's `Handle.comm_init`."""
    buf = C.create_string_buffer(nat.COMM_ID_BYTES)
    nat.check(nat.lib().dsmgp_comm_unique_id(C.cast(buf, C.c_void_p)))
    return buf.raw


class Handle:
    def __init__(self, x: np.ndarray, leaf_obs: Sequence[np.ndarray], y_centered: Sequence[np.ndarray],
                 leaf_mean: Sequence[float], leaf_kernel_id: Sequence[int], kernels: Sequence[KernelFunction],
                 tree: nat.FlatTree, *, as_written_grads: bool = True, keep_factors: bool = True, rank: int = 0,
                 world: int = 1, device: int = -1, strict_pd: bool = False, arena_bytes: int = 0):
        L = nat.lib()
        self._lib = L
        self._h = C.c_void_p()
        x = nat.colmajor(x)
        self.N, self.D = x.shape
        self.L = len(leaf_obs)
        leaf_ptr = np.zeros(self.L + 1, dtype=np.int64)
        leaf_ptr[1:] = np.cumsum([len(o) for o in leaf_obs])
        obs = np.ascontiguousarray(np.concatenate(leaf_obs), dtype=np.int64)
        yc = nat.f64(np.concatenate(y_centered))
        lm = nat.f64(leaf_mean)
        kid = np.ascontiguousarray(leaf_kernel_id, dtype=np.int32)
        kd = (nat.KernelDesc * len(kernels))(*[nat.KernelDesc(k.type, k.nparams) for k in kernels])
        opts = nat.Opts()
        L.dsmgp_default_opts(C.byref(opts))
        opts.as_written_grads = 1 if as_written_grads else 0
        opts.keep_factors = 1 if keep_factors else 0
        opts.rank, opts.world, opts.device = rank, world, device
        opts.strict_pd = 1 if strict_pd else 0
        opts.arena_bytes = int(arena_bytes)
        self.tree = tree
        self.kernels = list(kernels)
        self.leaf_ptr = leaf_ptr
        self.leaf_kernel_id = kid
        rc = L.dsmgp_create(nat.p_d(x), self.N, self.D, self.L, nat.p_i64(leaf_ptr), nat.p_i64(obs), nat.p_d(yc),
                            nat.p_d(lm), nat.p_i32(kid), kd, len(kernels), C.byref(tree.struct), C.byref(opts),
                            C.byref(self._h))
        if rc != nat.OK:
            self._h = C.c_void_p()
            nat.check(rc, None)
        self.nparams = int(L.dsmgp_nparams(self._h))
        self.n_nodes = int(L.dsmgp_n_nodes(self._h))
        self.row_width = int(L.dsmgp_row_width(self._h))
        self.world, self.rank = world, rank

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.dsmgp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        nat.check(rc, self._h)

    # ---- parameters
    def set_params(self, theta):
        th = nat.f64(theta)
        self._ck(self._lib.dsmgp_set_params(self._h, nat.p_d(th), th.size))

    def set_leaf_params(self, leaf: int, theta):
        th = nat.f64(theta)
        self._ck(self._lib.dsmgp_set_leaf_params(self._h, leaf, nat.p_d(th), th.size))

    def get_leaf_params(self, leaf: int) -> np.ndarray:
        n = self.kernels[int(self.leaf_kernel_id[leaf])].nparams
        th = np.zeros(n)
        self._ck(self._lib.dsmgp_get_leaf_params(self._h, leaf, nat.p_d(th), n))
        return th

    # ---- compute
    def fit(self, overlap=None, tau: float = 0.05):
        """dsmgp_fit(h, tau, overlap, info, seconds): overlap = the L x L matrix of getOverlap (shared Cholesky of fit!),
        or None for fit_naive!."""
        info = np.zeros(self.L, dtype=np.int32)
        sec = C.c_double(0)
        ov = None if overlap is None else nat.colmajor(np.asarray(overlap, dtype=np.float64))
        self._ck(self._lib.dsmgp_fit(self._h, float(tau), nat.p_d(ov), nat.p_i32(info), C.byref(sec)))
        return info, sec.value

    def set_sharing(self, overlap=None, tau: float = 0.05):
        """Store (or clear, overlap=None) the sharing plan used by every later fit / eval."""
        ov = None if overlap is None else nat.colmajor(np.asarray(overlap, dtype=np.float64))
        self._ck(self._lib.dsmgp_set_sharing(self._h, nat.p_d(ov), float(tau)))

    def get_sharing(self):
        """(kind, source, blocks) per leaf: 0 own factor / 1 identical to source / 2 continues behind copied block rows."""
        k = np.zeros(self.L, dtype=np.int32); s = np.zeros(self.L, dtype=np.int32); b = np.zeros(self.L, dtype=np.int32)
        self._ck(self._lib.dsmgp_get_sharing(self._h, nat.p_i32(k), nat.p_i32(s), nat.p_i32(b)))
        return k, s, b

    def comm_init(self, unique_id: bytes):
        """dsmgp_comm_init: attach this rank's handle to the NCCL communicator named by `unique_id` (collective call)."""
        buf = C.create_string_buffer(bytes(unique_id), nat.COMM_ID_BYTES)
        self._ck(self._lib.dsmgp_comm_init(self._h, C.cast(buf, C.c_void_p)))

    def lml(self) -> np.ndarray:
        out = np.zeros(self.n_nodes)
        self._ck(self._lib.dsmgp_lml(self._h, nat.p_d(out)))
        return out

    def grad(self, leaf_scale=None) -> np.ndarray:
        g = np.zeros(self.nparams)
        ls = None if leaf_scale is None else nat.f64(leaf_scale)
        self._ck(self._lib.dsmgp_grad(self._h, nat.p_d(ls), nat.p_d(g)))
        return g

    def eval(self, theta=None, leaf_scale=None, want_grad: bool = True, want_nodes: bool = False):
        th = None if theta is None else nat.f64(theta)
        ls = None if leaf_scale is None else nat.f64(leaf_scale)
        lml = C.c_double(0)
        g = np.zeros(self.nparams) if want_grad else None
        nodes = np.zeros(self.n_nodes) if want_nodes else None
        self._ck(self._lib.dsmgp_eval(self._h, nat.p_d(th), 0 if th is None else th.size, nat.p_d(ls), C.byref(lml),
                                      nat.p_d(g), nat.p_d(nodes)))
        return (lml.value, g, nodes) if want_nodes else (lml.value, g)

    def finetune_eval(self, anchors, thetas, overlap):
        """dsmgp_finetune_eval: (leaf_lml[G], grads[G, H], root_lml[G]) for the anchor experts under their own theta."""
        an = np.ascontiguousarray(anchors, dtype=np.int64)
        th = np.ascontiguousarray(thetas, dtype=np.float64).reshape(an.size, self.nparams)
        ov = nat.colmajor(np.asarray(overlap, dtype=np.float64))
        ll = np.zeros(an.size); gr = np.zeros((an.size, self.nparams)); rl = np.zeros(an.size)
        self._ck(self._lib.dsmgp_finetune_eval(self._h, an.size, nat.p_i64(an), nat.p_d(th), nat.p_d(ov), nat.p_d(ll),
                                               nat.p_d(gr), nat.p_d(rl)))
        return ll, gr, rl

    def train(self, theta, optimiser: int, eta: float, beta1: float, beta2: float, state_by_identity: bool,
              iterations: int, lam: float, earlystop: int):
        """dsmgp_train: the whole train! loop inside the library -> (final theta, LML trace)."""
        th = np.array(theta, dtype=np.float64)
        ell = np.zeros(iterations)
        nd = C.c_int64(0)
        self._ck(self._lib.dsmgp_train(self._h, optimiser, eta, beta1, beta2, 1 if state_by_identity else 0, iterations,
                                       lam, earlystop, nat.p_d(th), nat.p_d(ell), C.byref(nd)))
        return th, ell[:nd.value]

    def eval_local_dev(self, theta=None) -> int:
        th = None if theta is None else nat.f64(theta)
        ptr = C.c_void_p()
        self._ck(self._lib.dsmgp_eval_local_dev(self._h, nat.p_d(th), 0 if th is None else th.size, C.byref(ptr)))
        return int(ptr.value)

    def eval_finish_dev(self, leaf_scale=None):
        ls = None if leaf_scale is None else nat.f64(leaf_scale)
        lml = C.c_double(0)
        g = np.zeros(self.nparams)
        self._ck(self._lib.dsmgp_eval_finish_dev(self._h, nat.p_d(ls), C.byref(lml), nat.p_d(g), None))
        return lml.value, g

    def leaf_rows(self) -> np.ndarray:
        rows = np.zeros((self.L, self.row_width))
        self._ck(self._lib.dsmgp_leaf_rows(self._h, nat.p_d(rows)))
        return rows

    def leaf_owner(self) -> np.ndarray:
        o = np.zeros(self.L, dtype=np.int32)
        self._ck(self._lib.dsmgp_leaf_owner(self._h, nat.p_i32(o)))
        return o

    def leaf_info(self) -> np.ndarray:
        o = np.zeros(self.L, dtype=np.int32)
        self._ck(self._lib.dsmgp_leaf_info(self._h, nat.p_i32(o)))
        return o

    def update_weights(self):
        lw = np.zeros(max(int(self.tree.child_ptr[-1]), 1))
        z = C.c_double(0)
        self._ck(self._lib.dsmgp_update_weights(self._h, nat.p_d(lw), C.byref(z)))
        return lw, z.value

    def infer(self):
        lw = np.zeros(max(int(self.tree.child_ptr[-1]), 1))
        z = C.c_double(0)
        self._ck(self._lib.dsmgp_infer(self._h, nat.p_d(lw), C.byref(z)))
        return lw, z.value

    def reset_weights(self):
        lw = np.zeros(max(int(self.tree.child_ptr[-1]), 1))
        self._ck(self._lib.dsmgp_reset_weights(self._h, nat.p_d(lw)))
        return lw

    def predict(self, xtest, mode: int = nat.PREDICT_DSMGP):
        xt = nat.colmajor(np.asarray(xtest, dtype=np.float64).reshape(len(xtest), -1))
        T = xt.shape[0]
        mu = np.zeros(T); var = np.zeros(T)
        self._ck(self._lib.dsmgp_predict(self._h, nat.p_d(xt), T, mode, nat.p_d(mu), nat.p_d(var)))
        return mu, var

    def predict_local(self, xtest, mode: int = nat.PREDICT_DSMGP) -> np.ndarray:
        """dsmgp_predict_local: [mu | var] of the local experts in (leaf, routing order) layout, 0 elsewhere."""
        xt = nat.colmajor(np.asarray(xtest, dtype=np.float64).reshape(len(xtest), -1))
        tot = C.c_int64(0)
        self._ck(self._lib.dsmgp_predict_local(self._h, nat.p_d(xt), xt.shape[0], mode, None, C.byref(tot)))
        buf = np.zeros(2 * tot.value)
        self._ck(self._lib.dsmgp_predict_local(self._h, nat.p_d(xt), xt.shape[0], mode, nat.p_d(buf), C.byref(tot)))
        return buf

    def predict_finish(self, xtest, buf: np.ndarray, mode: int = nat.PREDICT_DSMGP):
        xt = nat.colmajor(np.asarray(xtest, dtype=np.float64).reshape(len(xtest), -1))
        T = xt.shape[0]
        mu = np.zeros(T); var = np.zeros(T)
        b = nat.f64(buf)
        self._ck(self._lib.dsmgp_predict_finish(self._h, nat.p_d(xt), T, mode, nat.p_d(b), nat.p_d(mu), nat.p_d(var)))
        return mu, var

    def leaf_predict(self, leaf: int, xtest):
        xt = nat.colmajor(np.asarray(xtest, dtype=np.float64).reshape(len(xtest), -1))
        T = xt.shape[0]
        mu = np.zeros(T); var = np.zeros(T)
        self._ck(self._lib.dsmgp_leaf_predict(self._h, leaf, nat.p_d(xt), T, nat.p_d(mu), nat.p_d(var)))
        return mu, var

    def leaf_alpha(self, leaf: int) -> np.ndarray:
        n = int(self.leaf_ptr[leaf + 1] - self.leaf_ptr[leaf])
        a = np.zeros(n)
        self._ck(self._lib.dsmgp_leaf_alpha(self._h, leaf, nat.p_d(a)))
        return a

    def leaf_factor(self, leaf: int) -> np.ndarray:
        n = int(self.leaf_ptr[leaf + 1] - self.leaf_ptr[leaf])
        F = np.zeros((n, n), order="F")
        self._ck(self._lib.dsmgp_leaf_factor(self._h, leaf, nat.p_d(F)))
        return F

    def timings(self) -> dict:
        t = nat.Timings()
        self._ck(self._lib.dsmgp_get_timings(self._h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in nat.Timings._fields_}

    def int8_info(self) -> dict:
        """The INT8 split path of the last evaluation (dsmgp_int8_info): op counts and CUDA-event times of its launches."""
        o = np.zeros(9)
        self._ck(self._lib.dsmgp_int8_info(self._h, nat.p_d(o), 9))
        return {"batches": int(o[0]), "slices": int(o[1]), "int8_ops": o[2], "fp64_equiv_flops": o[3], "gemm_ms": o[4],
                "slice_ms": o[5], "fp64_tile_ms": o[6], "pool_bytes": int(o[7]), "fp64_tile_flops": o[8]}

    def set_profiling(self, on: bool):
        self._ck(self._lib.dsmgp_set_profiling(self._h, 1 if on else 0))
