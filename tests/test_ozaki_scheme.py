"""CPU restatement of the arithmetic of the INT8 split path (csrc/ozaki.cuh): the slicing of FP64 rows into signed 7-bit
slices is error-free, the integer products accumulate exactly, and the Horner recombination of the slice groups reproduces
the FP64 product to the bound the design states (pairs with s + t >= S dropped: < 2^-7S of rowmax * colmax per term).
NumPy only -- documents and pins the scheme the CUDA kernels implement (tools/ozaki_proto.cu measures the same on the GPU)."""
import numpy as np
import pytest


def slice_rows(X, S):
    """x = 2^(e-6) * sum_s q_s 128^-s per row, |q_s| <= 64 (slice_kernel: frexp of the row maximum, rint, exact remainders)."""
    m = np.abs(X).max(axis=1)
    e = np.where(m > 0, np.frexp(m)[1], 0)
    v = X * np.ldexp(64.0, -e)[:, None]
    q = np.empty((S,) + X.shape, dtype=np.int64)
    for s in range(S):
        r = np.rint(v)
        q[s] = r.astype(np.int64)
        v = (v - r) * 128.0
    return q, np.ldexp(1.0, e - 6), v


def ozaki_product(A, B, S):
    qa, sa, _ = slice_rows(A, S)
    qb, sb, _ = slice_rows(B, S)
    acc = np.zeros((A.shape[0], B.shape[0]))
    for g in range(S - 1, -1, -1):                     # smallest weight first: acc = acc / 128 + G_g
        G = np.zeros((A.shape[0], B.shape[0]), dtype=np.int64)
        for s in range(g + 1):
            G += qa[s] @ qb[g - s].T                   # exact (int64 here; the kernel's int32 bound is checked below)
        assert np.abs(G).max() < 2 ** 31
        acc = acc * 0.0078125 + G.astype(np.float64)
    return acc * sa[:, None] * sb[None, :]


@pytest.mark.parametrize("S", [6, 7, 8])
def test_slices_are_exact_and_bounded(S):
    rng = np.random.default_rng(S)
    X = rng.standard_normal((64, 512)) * 10.0 ** rng.uniform(-6, 0, size=(64, 512))
    q, scale, rem = slice_rows(X, S)
    assert np.abs(q).max() <= 64
    rec = sum(q[s] * 128.0 ** -s for s in range(S)) * scale[:, None]
    # what is left after S slices is the exact remainder, below 2^-(7 S) of the row scale * 64
    bound = np.abs(X).max(axis=1) * 2.0 ** (1 - 7 * S)
    assert np.all(np.abs(X - rec) <= bound[:, None])
    assert np.all(np.abs(rem) <= 64.0)


@pytest.mark.parametrize("S,tol", [(6, 1e-10), (7, 1e-12), (8, 2e-15)])
def test_slice_products_reproduce_the_fp64_product(S, tol):
    rng = np.random.default_rng(10 + S)
    K = 2048
    A = rng.uniform(-1, 1, (96, K)) * 10.0 ** (-4 * rng.random((96, K)))
    B = rng.uniform(-1, 1, (80, K)) * 10.0 ** (-4 * rng.random((80, K)))
    ref = (A.astype(np.longdouble) @ B.astype(np.longdouble).T)
    C = ozaki_product(A, B, S)
    err = float(np.max(np.abs(C - ref)) / np.max(np.abs(ref)))
    fp64 = float(np.max(np.abs(A @ B.T - ref)) / np.max(np.abs(ref)))
    print(f"\n[ozaki scheme] S={S}: max error {err:.2e} of max|C| (NumPy FP64 matmul: {fp64:.2e})")
    assert err <= tol
    if S == 8:
        assert err <= 4 * fp64 + 1e-16


def test_int32_accumulators_cannot_overflow_at_the_supported_k():
    """|sum over K of q_s q_t| <= 4096 K per pair, at most 8 pairs share an accumulator: K <= 65,536 keeps it below 2^31."""
    assert 4096 * 8 * 65536 - 1 < 2 ** 31
    S, K = 8, 4096
    A = np.full((8, K), 1.0 - 2.0 ** -30); B = -A                  # every slice-0 digit at its extreme value 64 / -64
    qa, _, _ = slice_rows(A, S); qb, _, _ = slice_rows(B, S)
    assert abs(int((qa[0] @ qb[0].T)[0, 0])) == 64 * 64 * K
