"""Summarise DSMGP_OZAKI_TRACE / DSMGP_OZAKI_TRACE_SYRK (clock stamps of the persistent INT8 block-product kernel).
Per block: nk, then SM-clock stamps relative to the moment its first MMA round may start (TMEM released by the previous block):
[2] first operands landed, [3] round-0 MMAs issued, [4] round-0 accumulators complete, [5] round-1 MMAs issued, [6] round-1
complete, [7] TMEM read by the epilogue, [8] stores issued."""
import sys
import numpy as np
a = np.loadtxt(sys.argv[1])
nk = a[:, 0]
ghz = 1.965
us_all = a[:, 1:] / ghz / 1e3
print("blocks", len(a), "mean nk", nk.mean())
for lo, hi in ((1, 8), (8, 16), (16, 32), (32, 64), (64, 200)):
    m = (nk >= lo) & (nk < hi)
    if not m.any():
        continue
    u = us_all[m]
    ideal = nk[m].mean() * 36 * 72 / ghz / 1e3
    print(f"nk in [{lo},{hi}): {m.sum()} blocks, mean nk {nk[m].mean():.1f}; ideal MMA (S=8) {ideal:.1f} us")
    print("   first operands %.1f | r0 issued %.1f | r0 done %.1f | r1 issued %.1f | r1 done %.1f | TMEM drained %.1f | stores issued %.1f" % tuple(u[:, i].mean() for i in (1, 2, 3, 4, 5, 6, 7)))
    print("   occupancy of the MMA warp per block (start -> TMEM drained): %.1f us = %.0f %% of ideal" % (u[:, 6].mean(), 100 * ideal / u[:, 6].mean()))
