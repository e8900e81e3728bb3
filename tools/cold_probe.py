import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np
import bench
from deepstructuredmixtures_b200 import model as mdl
w = bench.WORKLOADS["cfg3"]
x, y, root, kern = bench.build_structure(w)
th = bench.thetas([kern.nparams], w["seed"])[0]
def cold(tag):
    for i in range(3):
        t0 = time.perf_counter(); m2 = mdl.DSMGP(root, x, y, [kern.copy()], -1.0); t1 = time.perf_counter()
        m2.handle.eval(th); t2 = time.perf_counter(); m2.close(); t3 = time.perf_counter()
        print(tag, "create %.1f eval %.1f close %.1f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3), flush=True)
cold("fresh")
model = mdl.DSMGP(root, x, y, [kern.copy()], -1.0)
model.handle.eval(th)
cold("with main model")
mdl.update_(model); mdl.predict(model, np.random.default_rng(0).random((40000, 8)))
cold("after predict")
