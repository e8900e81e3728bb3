"""The N>1 path on CPU: world_size-2 gloo.  Each rank owns the leaves dsmgp_host_shard gives it, fills their rows
(the oracle stands in for the device kernels here -- there is no GPU), zero elsewhere; one SUM all-reduce assembles
the table and every rank finishes the tree passes with dsmgp_host_tree_eval.  Result must equal the 1-rank value."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from test_abi import _oracle_root, _structure
    from conftest import orc
    from deepstructuredmixtures_b200 import distributed as dd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, y, root, ft, leaves, kernels = _structure(5)
    lp = np.concatenate([[0], np.cumsum([lf.nobs for lf in leaves])])
    owner = dd.shard_leaves(lp, world)
    theta = np.array([0.1, -0.2, 0.3, 0.1, -1.0])
    oroot = _oracle_root(x, y, ft, leaves, kernels)
    orc.setparams(oroot, theta)
    rows = torch.zeros((len(leaves), 1 + kernels[0].nparams), dtype=torch.float64)
    for lf in orc.getLeaves(oroot):
        if owner[lf.leaf_index] == rank:
            lf.gp.update_cholesky()
            rows[lf.leaf_index, 0] = lf.gp.mll()
            rows[lf.leaf_index, 1:] = torch.from_numpy(lf.gp.grad_mll())
    dd.allreduce_rows_(rows)
    node_lml, grad, lw, z = dd.host_tree_eval(ft, [0] * len(leaves), kernels, rows.numpy())
    if rank == 0:
        q.put((float(node_lml[ft.root]), grad.tolist(), int((owner == 0).sum()), int((owner == 1).sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_rank():
    import torch.multiprocessing as mp
    from test_abi import _oracle_root, _structure
    from conftest import orc
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    lml, grad, n0, n1 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert n0 > 0 and n1 > 0
    x, y, root, ft, leaves, kernels = _structure(5)
    oroot = _oracle_root(x, y, ft, leaves, kernels)
    o_lml, o_grad, _, _ = orc.evaluate(oroot, np.array([0.1, -0.2, 0.3, 0.1, -1.0]))
    assert abs(lml - o_lml) <= 1e-12 * abs(o_lml)
    assert np.allclose(grad, o_grad, rtol=1e-12, atol=1e-12)
