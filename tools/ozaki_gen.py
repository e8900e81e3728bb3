"""Operands of a real trailing update for tools/ozaki_proto.cu: rows of the Cholesky factor of an ArdSE Gram matrix.

Writes (M + N) x K doubles, row-major: A = L[K:K+M, :K], B = L[K+M:K+M+N, :K] of K = exp(-0.5 d^2 / l^2) + sigma^2 I on
uniform 8-D inputs (the shape of a cfg3 expert).  Usage: python tools/ozaki_gen.py out.bin M N K
"""
import sys
import numpy as np
from scipy.linalg import cholesky

out, M, N, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
n = K + M + N
rng = np.random.default_rng(7)
x = rng.uniform(size=(n, 8))
ell = np.exp(np.linspace(-0.3, 0.3, 8))
z = x / ell
sq = (z * z).sum(1)
G = np.exp(-0.5 * np.maximum(sq[:, None] + sq[None, :] - 2.0 * z @ z.T, 0.0))
G[np.diag_indices(n)] += np.exp(2 * -1.0)
L = cholesky(G, lower=True)
with open(out, "wb") as f:
    f.write(np.ascontiguousarray(L[K:K + M, :K]).tobytes())
    f.write(np.ascontiguousarray(L[K + M:K + M + N, :K]).tobytes())
print("factor", n, "row max range", np.abs(L[K:, :K]).max(1).min(), np.abs(L[K:, :K]).max(1).max())
