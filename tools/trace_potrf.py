"""Analyse a DSMGP_TRACE_FILE dump of potrf2 (per-task clock stamps; clock64 is per SM, so only same-SM differences
are meaningful).  Record = 8 int64: start, after C-stage, after contraction, after epilogue/factor, end,
sm | I<<16 | J<<32, and for diagonal tasks: end of the panel loop, end of the 16x16 inverses."""
import sys
import numpy as np

t = np.fromfile(sys.argv[1], dtype=np.int64).reshape(-1, 8)
t = t[t[:, 0] > 0]
clk = 1.965e3  # cycles per us
start, c1, c2, c3, end, pk, ta, tb = t.T
sm, I, J = pk & 0xFFFF, (pk >> 16) & 0xFFFF, (pk >> 32) & 0xFFFF
diag = I == J
tot = (end - start).sum() / clk / 1e3
print(f"tasks {len(t)} (diag {diag.sum()}); total task time {tot:.1f} SM-ms = {tot / 148:.2f} ms x 148 SMs")
for nm, a, b in (("first-chunk wait", start, c1), ("contraction", c1, c2), ("epilogue / factor", c2, c3), ("tail", c3, end)):
    print(f"  {nm:18s} panel {((b - a)[~diag]).sum() / clk / 1e3:8.1f}   diag {((b - a)[diag]).sum() / clk / 1e3:8.1f}  SM-ms")
print(f"  ideal contraction at the DMMA roof (16.68 us per 128-k block): panel {(J[~diag] * 16.68).sum() / 1e3:.1f}, "
      f"diag (lower triangle, 5/8) {(J[diag] * 16.68 * 0.625).sum() / 1e3:.1f} SM-ms")
sel = (~diag) & (J > 0)
per = ((c2 - c1) / clk)[sel] / J[sel]
print(f"  panel contraction per k-block: median {np.median(per):.2f} p10 {np.percentile(per, 10):.2f} p90 {np.percentile(per, 90):.2f} us")
d = t[diag]
print("  diagonal task (us, mean): contraction %.1f | panel loop %.1f | 16x16 inverses %.1f | W %.1f | tail %.1f" % (
    ((d[:, 2] - d[:, 1]) / clk).mean(), ((d[:, 6] - d[:, 2]) / clk).mean(), ((d[:, 7] - d[:, 6]) / clk).mean(),
    ((d[:, 3] - d[:, 7]) / clk).mean(), ((d[:, 4] - d[:, 3]) / clk).mean()))
gaps = 0
for s_ in np.unique(sm):
    q = t[sm == s_]; q = q[np.argsort(q[:, 0])]
    gaps += (q[1:, 0] - q[:-1, 4]).sum()
print(f"  gaps between tasks {gaps / clk / 1e3:.1f} SM-ms")
