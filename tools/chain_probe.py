"""Critical path of ONE expert (the strong-scaling bound): potrf / inverse phase times of a single GP of n points."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepstructuredmixtures_b200 as dsm
for n in (2560, 5008):
    rng = np.random.default_rng(1)
    x = rng.random((n, 8)); y = np.sin(x.sum(1)) + 0.1 * rng.standard_normal(n)
    gp = dsm.GaussianProcess(x, y, kernel=dsm.ArdSE(np.zeros(8), 0.0), logNoise=-1.0, run_cholesky=True)
    th = np.concatenate([np.zeros(8), [0.0, -1.0]])
    for _ in range(3):
        gp.model.handle.eval(th)
    t = gp.model.handle.timings()
    nb = (n + 127) // 128
    print(f"n={n} nb={nb}: potrf {t['potrf_ms']:.3f} ms ({t['potrf_ms']*1e3/nb:.1f} us per block column), "
          f"inverse {t['inverse_ms']:.3f} ms ({t['inverse_ms']*1e3/nb:.1f} us per block column), gram {t['gram_ms']:.3f}")
