"""Analyse a DSMGP_PTRACE_FILE dump of predict3 (per-task cycle counts, device path): where the SM time of a prediction goes.
Record = 8 int64: start clock, contraction, staging, kernel values, TRSM epilogue, store + signal, end clock, row blocks."""
import sys
import numpy as np
t = np.fromfile(sys.argv[1], dtype=np.int64).reshape(-1, 8)
t = t[t[:, 0] > 0]
clk = 1.965e3
tot = (t[:, 6] - t[:, 0]).sum() / clk / 1e3
names = ["contraction", "staging", "kernel values", "TRSM epilogue", "store+signal"]
print(f"tasks {len(t)}, total task time {tot:.1f} SM-ms = {tot / 148:.2f} ms x 148 SMs; row blocks per task mean {t[:, 7].mean():.1f}")
for i, nm in enumerate(names):
    v = t[:, 1 + i].sum() / clk / 1e3
    print(f"  {nm:14s} {v:9.1f} SM-ms  {100 * v / tot:5.1f} %   per row block {t[:, 1 + i].sum() / t[:, 7].sum() / clk:7.2f} us")
nb = t[:, 7]
ideal = (nb * (nb - 1) / 2 * 16.68).sum() / 1e3
print(f"  ideal contraction at the DMMA roof: {ideal:.1f} SM-ms ({100 * ideal / tot:.1f} % of task time); measured per k-block "
      f"{t[:, 1].sum() / (nb * (nb - 1) / 2).sum() / clk:.2f} us")
rest = tot - t[:, 1:6].sum() / clk / 1e3
print(f"  outside the phases (task setup, final reductions): {rest:.1f} SM-ms")
