"""torchrun check of the leaf-sharded path over NCCL: evaluate_distributed + update + predict_distributed against a
single-handle model on rank 0.  usage: torchrun --nproc-per-node N tools/dist_predict_check.py [workload]"""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import deepstructuredmixtures_b200 as dsm
from deepstructuredmixtures_b200 import model as mdl, distributed as dd
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
w = bench.WORKLOADS[wl]
x, y, root, kern = bench.build_structure(w)
klist = kern if isinstance(kern, list) else [kern]
th = bench.thetas([k.nparams for k in klist], w["seed"])[1]
m = mdl.DSMGP(root, x, y, [k.copy() for k in klist], -1.0, rank=rank, world=world, device=local)
lml, g = dd.evaluate_distributed(m, th)
z = dd.update_distributed(m)
xt = np.random.default_rng(9).random((20000, w["D"]))
dd.predict_distributed(m, xt)
dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
mu, var = dd.predict_distributed(m, xt)
torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
if rank == 0:
    s = mdl.DSMGP(root, x, y, [k.copy() for k in klist], -1.0, device=local)
    lml0, g0 = s.handle.eval(th); z0 = dsm.update_(s)
    dsm.predict(s, xt); t0 = time.perf_counter(); mu0, var0 = dsm.predict(s, xt); dt0 = time.perf_counter() - t0
    ok = (abs(lml - lml0) <= 1e-12 * abs(lml0) and np.allclose(g, g0, rtol=1e-12) and abs(z - z0) <= 1e-12 * abs(z0)
          and np.allclose(mu, mu0, rtol=1e-10, atol=1e-10 * np.max(np.abs(mu0))) and np.allclose(var, var0, rtol=1e-10))
    print(f"{wl} x{world}: sharded == single handle: {ok}; predict 20000 points {dt * 1e3:.1f} ms on {world} GPUs vs {dt0 * 1e3:.1f} ms on 1", flush=True)
    s.close()
m.close()
dist.destroy_process_group()
