"""Analyse a DSMGP_TRACE_FILE dump of potrf2 (per-task clock stamps)."""
import sys, numpy as np
t = np.fromfile(sys.argv[1], dtype=np.int64).reshape(-1, 8)
t = t[t[:, 0] > 0]
start, c1, c2, c3, end, sm, I, J = t.T
diag = I == J
clk = 1.965e3  # cycles per us
print("tasks", len(t), "diag", diag.sum())
for name, sel in (("panel", ~diag), ("diag", diag)):
    s = t[sel]
    tot = (s[:, 4] - s[:, 0]) / clk
    print(f"{name}: n={len(s)} mean total {tot.mean():.1f} us; C-stage wait {((s[:,1]-s[:,0])/clk).mean():.2f}; main {((s[:,2]-s[:,1])/clk).mean():.1f}; "
          f"epi/factor {((s[:,3]-s[:,2])/clk).mean():.2f}; tail {((s[:,4]-s[:,3])/clk).mean():.2f}")
    Jm = np.maximum(s[:, 7], 1)
    print(f"   main per k-block: {(((s[:,2]-s[:,1])/clk)[s[:,7]>0] / s[:,7][s[:,7]>0]).mean():.2f} us (ideal 16.7 @peak)")
# per-SM busy fraction and gaps between consecutive tasks on the same SM
span = (end.max() - start.min()) / clk
busy = 0
gaps = []
for s_ in np.unique(sm):
    q = t[sm == s_]; q = q[np.argsort(q[:, 0])]
    busy += ((q[:, 4] - q[:, 0]) / clk).sum()
    gaps.extend(((q[1:, 0] - q[:-1, 4]) / clk).tolist())
print(f"kernel span {span/1e3:.2f} ms; mean SM busy {busy/len(np.unique(sm))/span*100:.1f}%; mean gap between tasks {np.mean(gaps):.2f} us")
last = np.array([t[sm == s_][:, 4].max() for s_ in np.unique(sm)])
print(f"SM finish spread: min {(last.min()-start.min())/clk/1e3:.2f} ms, max {(last.max()-start.min())/clk/1e3:.2f} ms")
