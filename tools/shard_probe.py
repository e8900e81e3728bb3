"""Emulate every rank of an N-way leaf shard on ONE GPU (the handle takes rank/world): per-rank phase times.
usage: shard_probe.py [world] [workload]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from deepstructuredmixtures_b200 import model as mdl
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
wl = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
only = int(sys.argv[3]) if len(sys.argv) > 3 else -1
w = bench.WORKLOADS[wl]
x, y, root, kern = bench.build_structure(w)
klist = kern if isinstance(kern, list) else [kern]
th = bench.thetas([k.nparams for k in klist], w["seed"])[0]
worst = 0.0
for r in range(world):
    if only >= 0 and r != only:
        continue
    m = mdl.DSMGP(root, x, y, [k.copy() for k in klist], -1.0, rank=r, world=world)
    H = m.handle
    own = np.where(H.leaf_owner() == r)[0]
    sizes = sorted((int(H.leaf_ptr[l + 1] - H.leaf_ptr[l]) for l in own), reverse=True)
    for _ in range(3):
        H.eval_local_dev(th)
        import ctypes
    t = H.timings()
    fl = t["potrf_flops"]
    print(f"rank {r}: {len(own)} experts (max n {sizes[0]}), gram {t['gram_ms']:.2f} potrf {t['potrf_ms']:.2f} "
          f"({fl / t['potrf_ms'] * 1e-9:.1f} TF) inverse {t['inverse_ms']:.2f} total {t['total_ms']:.2f} ms", flush=True)
    worst = max(worst, t["total_ms"])
    m.close()
print(f"slowest rank {worst:.2f} ms -> {1e3 / worst:.1f} evals/s at {world} GPUs (without the all-reduce)")
