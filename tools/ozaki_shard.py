"""Per-shard evaluation time on ONE GPU (rank r of `world` emulated: the handle owns only that rank's experts, no collective):
FP64 pipeline (fused launch on small shards) against the INT8 split path.  Usage: python tools/ozaki_shard.py [world]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def shard_ms(world, rank, ozaki):
    os.environ["DSMGP_OZAKI"] = "1" if ozaki else "0"
    if ozaki:
        os.environ["DSMGP_FUSED_EVAL"] = "0"
    else:
        os.environ.pop("DSMGP_FUSED_EVAL", None)
    from deepstructuredmixtures_b200 import model as mdl
    w = bench.WORKLOADS["cfg3"]
    x, y, root, kern = bench.build_structure(w, device=True)
    model = mdl.DSMGP(root, x, y, [kern.copy()], -1.0, rank=rank, world=world, keep_factors=True)
    H = model.handle
    th = bench.thetas([kern.nparams], w["seed"])[0]
    for _ in range(3):
        H.eval_local_dev(th)
    tot = 0.0
    for _ in range(5):
        H.eval_local_dev(th)
        tot += H.timings()["total_ms"] / 5
    H.close()
    return tot


world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ranks = [int(sys.argv[2])] if len(sys.argv) > 2 else range(world)
for r in ranks:
    a = shard_ms(world, r, False)
    b = shard_ms(world, r, True)
    print(f"world {world} rank {r}: FP64 {a:.3f} ms   INT8 split {b:.3f} ms")
