"""A/B of the split inverse on the INT8 tensor cores (DSMGP_OZAKI=1) against the FP64 tile pipeline on a bench workload.

Builds the model twice in one process (the plan is made at create), evaluates the same theta, compares the per-expert rows
[lml, gradient...], the model LML / gradient and alpha of the largest expert, and prints the phase timings.
Usage: python tools/ozaki_ab.py [workload] [--mathematical]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def run(workload, ozaki, math, steps=3):
    os.environ["DSMGP_OZAKI"] = "1" if ozaki else "0"
    from deepstructuredmixtures_b200 import model as mdl
    w = bench.WORKLOADS[workload]
    x, y, root, kern = bench.build_structure(w, device=True)
    klist = kern if isinstance(kern, list) else [kern]
    model = mdl.DSMGP(root, x, y, [k.copy() for k in klist], -1.0, keep_factors=True, as_written_grads=not math)
    H = model.handle
    ths = bench.thetas([k.nparams for k in klist], w["seed"])
    for i in range(3):
        H.eval(ths[0])
    tm = None
    acc = {}
    for i in range(steps):
        out = H.eval(ths[0])
        tm = H.timings()
        for k in ("gram_ms", "potrf_ms", "inverse_ms", "grad_ms", "total_ms"):
            acc[k] = acc.get(k, 0.0) + tm[k] / steps
    rows = H.leaf_rows().copy()
    sizes = np.array([lf.nobs for lf in model.leaves])
    big = int(np.argmax(sizes))
    alpha = H.leaf_alpha(big).copy()
    res = {"out": out, "rows": rows, "alpha": alpha, "tm": acc, "launches": tm["launches"], "sizes": sizes}
    model.close() if hasattr(model, "close") else H.close()
    return res


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "cfg3"
    math = "--mathematical" in sys.argv
    a = run(workload, False, math)
    b = run(workload, True, math)
    la, ga = a["out"][0], np.asarray(a["out"][1])
    lb, gb = b["out"][0], np.asarray(b["out"][1])
    print("phases FP64 :", {k: round(v, 3) for k, v in a["tm"].items()}, "launches", a["launches"])
    print("phases INT8 :", {k: round(v, 3) for k, v in b["tm"].items()}, "launches", b["launches"])
    print("model lml rel diff %.3e" % (abs(la - lb) / abs(la)))
    print("model grad rel diff %.3e (scale %.3e)" % (np.abs(ga - gb).max() / np.abs(ga).max(), np.abs(ga).max()))
    ra, rb = a["rows"], b["rows"]
    d = np.abs(ra - rb)
    sc = np.maximum(np.abs(ra), 1e-300)
    print("rows: lml col max rel %.3e; grad cols max abs %.3e / max |g| %.3e; worst row rel (vs row max) %.3e" % (
        (d[:, 0] / sc[:, 0]).max(), d[:, 1:].max(), np.abs(ra[:, 1:]).max(), (d[:, 1:].max(1) / np.maximum(np.abs(ra[:, 1:]).max(1), 1e-300)).max()))
    print("alpha (largest expert, n = %d) rel diff %.3e" % (a["sizes"].max(), np.abs(a["alpha"] - b["alpha"]).max() / np.abs(a["alpha"]).max()))


if __name__ == "__main__":
    main()
