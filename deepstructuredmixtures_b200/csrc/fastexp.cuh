// Table-driven FP64 exp for non-positive arguments (the only kind the SE kernels produce: -0.5 r^2 / l^2).
//
//   x = (64 m + j) ln2/64 + r,  |r| <= ln2/128   =>   exp(x) = 2^m * 2^(j/64) * exp(r)
// 2^(j/64) comes from a 64-entry table in shared memory, exp(r) - 1 from a degree-5 polynomial (next term r^6/720 <=
// 3.5e-17), the scaling by 2^m is an integer add on the exponent field.  10 FP64 instructions + 1 shared load instead
// of the ~21 FP64 issue slots of the library exp() (measured: 818 Gexp/s against 17 T DFMA/s), max relative error
// 2.3e-16 against mpmath on 1e6 points (tests/test_oracle.py restates it in NumPy).  Arguments below -708 (results
// below 3e-308, subnormal) return 0; positive arguments are NOT supported.
#pragma once
#include "common.cuh"

namespace dsm {

constexpr int EXPTAB_N = 64;
__device__ const double g_exptab[EXPTAB_N] = {
  1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
  1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
  1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
  1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
  1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
  1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
  1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
  1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
  1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
  1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
  1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
  1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
  1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
  1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
  1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
  1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951};

// every thread of the block helps; the caller synchronises before the first use
__device__ __forceinline__ void exptab_load(double* sT) {
  for (int i = threadIdx.x; i < EXPTAB_N; i += blockDim.x) sT[i] = g_exptab[i];
}

__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ sT) {
  const double MAGIC = 6755399441055744.0;                         // 1.5 * 2^52: the low word of (t + MAGIC) is rint(t)
  const double t = fma(x, 92.33248261689366, MAGIC);               // 64 / ln2
  const int k = __double2loint(t);
  const double kf = t - MAGIC;
  double r = fma(kf, -0.01083042469326756, x);                     // ln2/64, upper 32 bits (k has 17 bits: exact product)
  r = fma(kf, -2.9815858269852933e-12, r);                         // ln2/64, remainder
  double p = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
  p = fma(r, p, 1.6666666666666666e-01);
  p = fma(r, p, 0.5);
  p = fma(r, p, 1.0);
  p = p * r;                                                       // exp(r) - 1
  const double tj = sT[k & (EXPTAB_N - 1)];
  const double v = fma(tj, p, tj);
  const int hi = __double2hiint(v) + ((k >> 6) << 20);             // * 2^m (v in [1, 2): no carry into the sign)
  return x < -708.0 ? 0.0 : __hiloint2double(hi, __double2loint(v));
}

}  // namespace dsm
