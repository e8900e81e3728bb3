"""Race-detection proxy (compute-sanitizer is not available on this pool): every kernel of the path repeated many times on
the same inputs must give bit-identical results (all reductions are ordered; cross-CTA flags carry the dependencies)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepstructuredmixtures_b200 as dsm
from deepstructuredmixtures_b200 import model as mdl, structure as st
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(5)
bad = 0
for (N, D, kern, M, mathematical) in ((9000, 8, dsm.ArdSE(np.zeros(8), 0.0), 150, False), (6000, 3, dsm.IsoSE(0.0, 0.0), 120, False),
                                      (5000, 12, dsm.ArdSE(np.zeros(12), 0.0), 200, True)):
    x = rng.random((N, D)); y = np.sin(2 * np.pi * x @ rng.standard_normal(D)) + 0.1 * rng.standard_normal(N)
    m = dsm.buildDSMGP(x, y, 3, 4, M=M, kernel=kern, logNoise=-1.0, rng=7, as_written_grads=not mathematical)
    th = np.concatenate([0.2 * rng.standard_normal(kern.logl.size), [0.1, -1.0]])
    xt = rng.random((3000, D)); xs = rng.random((5, D))
    ref = None
    for r in range(reps):
        lml, g = m.handle.eval(th)
        rows = m.handle.leaf_rows().copy()
        dsm.update_(m)
        mu, var = dsm.predict(m, xt)
        mu1, var1 = dsm.predict(m, xs)
        info, _ = m.handle.fit()
        al = m.handle.leaf_alpha(0).copy()
        cur = (lml, g.copy(), rows, mu, var, mu1, var1, al)
        if ref is None:
            ref = cur
        else:
            same = all(np.array_equal(a, b) for a, b in zip(ref, cur))
            if not same:
                bad += 1
                print("MISMATCH at repetition", r, [bool(np.array_equal(a, b)) for a, b in zip(ref, cur)], flush=True)
    print(type(kern).__name__, "D", D, "experts", len(m.leaves), "n", min(l.nobs for l in m.leaves), max(l.nobs for l in m.leaves),
          "reps", reps, "lml", ref[0], flush=True)
    m.close()
if len(sys.argv) > 2 and sys.argv[2] == "big":      # cfg3 at full size: large tiles, 40k-point predict (one task per expert and block)
    import bench
    w = bench.WORKLOADS["cfg3"]
    x, y, root, kern = bench.build_structure(w)
    m = mdl.DSMGP(root, x, y, [kern.copy()], -1.0)
    th = bench.thetas([kern.nparams], w["seed"])[2]
    xt = rng.random((40000, 8))
    ref = None
    for r in range(8):
        lml, g = m.handle.eval(th)
        dsm.update_(m)
        mu, var = dsm.predict(m, xt)
        cur = (lml, g.copy(), mu, var)
        if ref is None:
            ref = cur
        elif not all(np.array_equal(a, b) for a, b in zip(ref, cur)):
            bad += 1
            print("MISMATCH (cfg3) at repetition", r, flush=True)
    print("cfg3 full size: 8 repetitions, lml", ref[0], flush=True)
    m.close()
print("mismatches:", bad)
sys.exit(1 if bad else 0)
