"""Summarise DSMGP_OZAKI_TRACE (clock stamps of the INT8 block-product kernel): per tile nk + 7 stamps relative to the CTA start:
setup done, first operands landed, round-0 MMAs issued, round-0 accumulators complete, round-0 epilogue done, round-1 complete, round-1 epilogue done."""
import sys
import numpy as np
a = np.loadtxt(sys.argv[1])
nk = a[:, 0]
ghz = 1.965
names = ["setup", "first_full", "r0_issued", "r0_done", "r0_epi", "r1_done", "r1_epi"]
print("tiles", len(a), "mean nk", nk.mean())
for lo, hi in ((1, 8), (8, 16), (16, 32), (32, 64), (64, 200)):
    m = (nk >= lo) & (nk < hi)
    if not m.any():
        continue
    s = a[m]
    us = s[:, 1:] / ghz / 1e3
    d = np.diff(np.concatenate([np.zeros((len(us), 1)), us], axis=1), axis=1)
    print(f"nk in [{lo},{hi}): {m.sum()} tiles, mean nk {s[:,0].mean():.1f}; total {us[:, -1].mean():.1f} us; ideal MMA {s[:,0].mean()*36*72/ghz/1e3:.1f} us")
    print("   cumulative us:", " ".join(f"{n}={v:.1f}" for n, v in zip(names, us.mean(0))))
    print("   r0 mma span %.1f  r0 epilogue %.1f  r1 mma span %.1f  r1 epilogue %.1f" % ((us[:, 3] - us[:, 1]).mean(), (us[:, 4] - us[:, 3]).mean(), (us[:, 5] - us[:, 4]).mean(), (us[:, 6] - us[:, 5]).mean()))
