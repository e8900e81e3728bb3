"""One-process-per-GPU plumbing: leaf sharding and the single collective of an evaluation.

Leaves are independent given theta (fit.jl:88-119, 306-311), so they shard across ranks with no data-path
exchange; the only exchange step is the table of per-leaf rows [mll(gp), ∇mll(gp)...] (L x (1+H) doubles) that the
O(L) tree passes (optimize.jl:27-89) need.  Each rank fills the rows of its own leaves and leaves the others 0, so a
SUM all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests) assembles the table on every rank, which then
finishes the passes redundantly.  `torch.distributed` is plumbing only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _native as nat


def shard_leaves(leaf_ptr: Sequence[int], world: int) -> np.ndarray:
    """LPT bin packing of leaves by n^3 (dsmgp_host_shard; identical to what dsmgp_create does internally)."""
    lp = np.ascontiguousarray(leaf_ptr, dtype=np.int64)
    owner = np.zeros(lp.size - 1, dtype=np.int32)
    nat.check(nat.lib().dsmgp_host_shard(lp.size - 1, nat.p_i64(lp), int(world), nat.p_i32(owner)))
    return owner


def host_tree_eval(flat: nat.FlatTree, leaf_kernel_id: Sequence[int], kernels, rows: np.ndarray,
                   leaf_scale: Optional[np.ndarray] = None):
    """Tree passes over an assembled row table (dsmgp_host_tree_eval): returns (node_lml, grad, sum_logweights, z)."""
    rows = nat.f64(rows)
    L, rw = rows.shape
    kid = np.ascontiguousarray(leaf_kernel_id, dtype=np.int32)
    kd = (nat.KernelDesc * len(kernels))(*[nat.KernelDesc(k.type, k.nparams) for k in kernels])
    H = sum(k.nparams for k in kernels)
    node_lml = np.zeros(len(flat.node_type))
    grad = np.zeros(H)
    lw = np.zeros(max(int(flat.child_ptr[-1]), 1))
    z = C.c_double(0)
    ls = None if leaf_scale is None else nat.f64(leaf_scale)
    nat.check(nat.lib().dsmgp_host_tree_eval(C.byref(flat.struct), L, nat.p_i32(kid), kd, len(kernels), nat.p_d(rows), rw,
                                             nat.p_d(ls), nat.p_d(node_lml), nat.p_d(grad), nat.p_d(lw), C.byref(z)))
    return node_lml, grad, lw, z.value


class _DevPtr:
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}


def allreduce_rows_(rows) -> None:
    """In-place SUM all-reduce of the row table (a torch tensor on the rank's device, or on the CPU with gloo)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(rows, op=dist.ReduceOp.SUM)


def init_library_comm(model) -> None:
    """Attach the model's handle to an NCCL communicator owned by libdsmgp (`dsmgp_comm_init`): afterwards `evaluate`,
    `fit_`, `update_` and `predict` work on the sharded model exactly as on a single GPU -- the row table is all-reduced
    inside the library.  The 128-byte NCCL id is created by rank 0 (`dsmgp_comm_unique_id`) and distributed through the
    default process group's store (plumbing only; a Julia host would use MPI.jl or a file)."""
    import torch.distributed as dist
    from ._handle import comm_unique_id
    ids = [comm_unique_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    model.handle.comm_init(ids[0])


def evaluate_distributed(model, theta, leaf_scale=None) -> Tuple[float, np.ndarray]:
    """One LML+gradient evaluation of a model whose leaves are sharded over the ranks of the default process group
    (the model must have been created with rank=dist.get_rank(), world=dist.get_world_size())."""
    import torch
    H = model.handle
    ptr = H.eval_local_dev(theta)
    rows = torch.as_tensor(_DevPtr(ptr, H.L * H.row_width), device=torch.device("cuda", torch.cuda.current_device()))
    allreduce_rows_(rows)
    torch.cuda.synchronize()
    return H.eval_finish_dev(leaf_scale)


def update_distributed(model) -> float:
    """update!(model) on a sharded model: the row table is complete on every rank after `evaluate_distributed`."""
    from .model import update_
    return update_(model)


def predict_distributed(model, xtest) -> Tuple[np.ndarray, np.ndarray]:
    """predict(model, x) of a model whose experts are sharded over the ranks of the default process group: every rank
    predicts its own experts (`dsmgp_predict_local`), ONE SUM all-reduce assembles the per-(expert, point) table, every
    rank mixes (`dsmgp_predict_finish`).  Call after `evaluate_distributed` (or `fit_`) and `update_distributed`."""
    import torch
    import torch.distributed as dist
    H = model.handle
    buf = H.predict_local(xtest, model.predict_mode)
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.from_numpy(buf)
        if dist.get_backend() == "nccl":
            t = t.cuda()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            buf = t.cpu().numpy()
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return H.predict_finish(xtest, buf, model.predict_mode)
