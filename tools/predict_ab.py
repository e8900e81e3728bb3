"""A/B of the predict path: wall and kernel time of dsmgp_predict on the cfg3 model (40,000 test points) for the library named by
DSMGP_LIB_PATH, host path vs device path.  usage: python tools/predict_ab.py [T]"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from deepstructuredmixtures_b200 import model as mdl

T = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
w = bench.WORKLOADS[os.environ.get("WL", "cfg3")]
x, y, root, kern = bench.build_structure(w)
klist = kern if isinstance(kern, list) else [kern]
model = mdl.DSMGP(root, x, y, [k.copy() for k in klist], -1.0)
ths = bench.thetas([k.nparams for k in klist], w["seed"])
model.handle.eval(ths[1]); mdl.update_(model)
xt = np.random.default_rng(77).random((T, w["D"]))
out = {"lib": os.environ.get("DSMGP_LIB_PATH", "default"), "T": T}
ref = None
for path in ("0", "1"):
    os.environ["DSMGP_PREDICT_DEVICE"] = path
    mdl.predict(model, xt)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); mu, var = mdl.predict(model, xt); ts.append(time.perf_counter() - t0)
    tm = model.handle.timings()
    out["device" if path == "1" else "host"] = {"wall_ms": 1e3 * min(ts), "kernel_ms": tm["predict_ms"],
                                                "kernel_tflops": tm["predict_flops"] / (tm["predict_ms"] * 1e-3) * 1e-12}
    if ref is None:
        ref = (mu, var)
    else:
        out["max_rel_diff"] = float(max(np.max(np.abs(mu - ref[0])) / np.max(np.abs(ref[0])), np.max(np.abs(var - ref[1]) / np.abs(ref[1]))))
print(json.dumps(out))
