"""Time the public predict call vs its kernel on a bench workload.  usage: predict_probe.py [workload]"""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from deepstructuredmixtures_b200 import model as mdl
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
w = bench.WORKLOADS[wl]
x, y, root, kern = bench.build_structure(w)
klist = kern if isinstance(kern, list) else [kern]
model = mdl.DSMGP(root, x, y, [k.copy() for k in klist], -1.0)
H = model.handle
th = bench.thetas([k.nparams for k in klist], w["seed"])[0]
H.eval(th)
mdl.update_(model)
rng = np.random.default_rng(77)
for T in (1, 1, 1, 1, 16, 16, 256, 256, 40000, 40000, 1, 1):
    xt = rng.random((T, w["D"]))
    t0 = time.perf_counter(); mu, var = mdl.predict(model, xt); dt = time.perf_counter() - t0
    tm = H.timings()
    print(wl, "T", T, "wall ms %.2f kernel ms %.3f  kernel GB/s %.0f TF %.2f" % (dt * 1e3, tm["predict_ms"], tm["predict_bytes"] / tm["predict_ms"] * 1e-6, tm["predict_flops"] / tm["predict_ms"] * 1e-9), flush=True)
