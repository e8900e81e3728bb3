// Routing of test points through the region graph and mixing of the experts' predictions, on the device.
//
// Replaces, for large batches of test points, the host recursion of predict(model, x):
//   getchild                 common.jl:101-122   child k of a split node iff s_{k-1} < x_d <= s_k (k = 1: x_d <= s_1)
//   _minpredict              common.jl:151-173   min over the experts a point reaches (sum: all children, split: routed child)
//   _predict / predict       common.jl:134-143, 181-196, 275-302   log-space moments, lse with the sum nodes' posterior weights
//   lse                      common.jl:309-313   max-shifted
//   _predictPoE / gPoE / rBCM  common.jl:145-149, 198-241           precision-weighted products (every expert sees every point)
// One thread per test point walks the (small) flattened tree with an explicit stack.  The experts a point reaches are
// enumerated in depth-first child order by all three kernels, so the r-th expert of a point is the same in each of them.
#pragma once
#include "common.cuh"
#include "route_args.h"

namespace dsm {

__device__ __forceinline__ int dev_getchild(const DevTree& t, int node, const double* x, int64_t T, int64_t p) {
  const double xv = x[(int64_t)t.split_dim[node] * T + p];
  const int K = t.child_ptr[node + 1] - t.child_ptr[node];
  const double* s = t.split_val + t.split_ptr[node];
  for (int k = 0; k < K; k++) {
    const bool ok = (k == 0) ? (xv <= s[0]) : ((xv <= s[k]) && (xv > s[k - 1]));
    if (ok) return k;
  }
  return -1;
}

// FILL = false: count the points of every expert.  FILL = true: assign positions, write pidx / reach.
template <bool FILL>
__global__ void __launch_bounds__(256) route_kernel(RouteArgs a) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= a.T) return;
  if (!FILL) {
    for (int d = 0; d < a.D; d++) if (!isfinite(a.xtest[(int64_t)d * a.T + p])) { atomicExch(a.err, 1); return; }
  }
  int stack[ROUTE_STACK];
  int sp = 0, r = 0;
  stack[sp++] = a.t.root;
  while (sp > 0) {
    const int node = stack[--sp];
    const int ty = a.t.type[node];
    if (ty == 0) {
      const int l = a.t.leaf_of_node[node];
      if (FILL) {
        const int pos = a.ooff[l] + atomicAdd(a.fill + l, 1);
        a.pidx[pos] = (int)p;
        a.reach[p * a.R + r] = pos;
      } else {
        atomicAdd(a.cnt + l, 1);
      }
      r++;
      continue;
    }
    const int c0 = a.t.child_ptr[node], K = a.t.child_ptr[node + 1] - c0;
    if (ty == 1 && !a.poe) {
      const int k = dev_getchild(a.t, node, a.xtest, a.T, p);
      if (k < 0) { atomicExch(a.err, 2); return; }
      stack[sp++] = a.t.child_idx[c0 + k];
    } else {
      for (int k = K - 1; k >= 0; k--) stack[sp++] = a.t.child_idx[c0 + k];     // reversed: popped in child order
    }
  }
}

__global__ void __launch_bounds__(128) mix_kernel(MixArgs a) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= a.T) return;
  const DevTree& t = a.t;
  const int* rc = a.reach + p * a.R;
  if (a.mode == 0) {
    // ---- DSMGP: predict(root) common.jl:243-254 (split), :294-302 (sum), :175-179 (leaf)
    double mumin = INFINITY;                                  // _minpredict: min over every expert the point reaches
    for (int r = 0; r < a.R; r++) { const int o = rc[r]; if (o < 0) break; mumin = fmin(mumin, a.lmu[o]); }
    mumin -= 1.0;                                             // _predict(node, x, mumin .- 1)
    int fnode[MIX_FRAMES], fk[MIX_FRAMES];
    double fval[MIX_FRAMES][MIX_KMAX][3];
    int nf = 0, r = 0, node = t.root;
    bool any_sum = false;
    double v0 = 0, v1 = 0, v2 = 0;
    for (;;) {
      // descend to the next expert
      for (;;) {
        const int ty = t.type[node];
        if (ty == 0) break;
        const int c0 = t.child_ptr[node];
        if (ty == 1) { node = t.child_idx[c0 + dev_getchild(t, node, a.xtest, a.T, p)]; continue; }
        any_sum = true;
        fnode[nf] = node; fk[nf] = 0; nf++;
        node = t.child_idx[c0];
      }
      {   // leaf: _predict(node::GPNode) common.jl:134-143
        const int o = rc[r++];
        const double m = a.lmu[o];
        double s2 = a.lvar[o];
        if (s2 <= 0) s2 = 1e-8;
        v0 = log(m - mumin); v1 = log(m * m); v2 = log(s2);
      }
      // return to the enclosing sum nodes
      bool done = false;
      for (;;) {
        if (nf == 0) { done = true; break; }
        const int f = nf - 1, nd = fnode[f], c0 = t.child_ptr[nd], K = t.child_ptr[nd + 1] - c0;
        const double lw = a.logw[c0 + fk[f]];
        fval[f][fk[f]][0] = v0 + lw; fval[f][fk[f]][1] = v1 + lw; fval[f][fk[f]][2] = v2 + lw;      // common.jl:284-286
        fk[f]++;
        if (fk[f] < K) { node = t.child_idx[c0 + fk[f]]; break; }
        double out[3];
#pragma unroll
        for (int q = 0; q < 3; q++) {                         // lse common.jl:309-313
          double m = -INFINITY;
          for (int k = 0; k < K; k++) m = fmax(m, fval[f][k][q]);
          double s = 0.0;
          for (int k = 0; k < K; k++) s += exp(fval[f][k][q] - m);
          out[q] = log(s) + m;
        }
        v0 = out[0]; v1 = out[1]; v2 = out[2];
        nf--;
      }
      if (done) break;
    }
    const double m = exp(v0) + mumin;
    a.mu[p] = m;
    a.var[p] = any_sum ? exp(v2) + (exp(v1) - m * m) : exp(v2);
    return;
  }
  // ---- PoE family: every expert predicts the point; split nodes combine precisions (common.jl:198-208)
  const int root = t.root, rc0 = t.child_ptr[root], RK = t.child_ptr[root + 1] - rc0;
  double s_prior = 0.0, C = 0.0, M = 0.0, Tt = 0.0;
  if (a.mode == 3) {                                          // _predictrBCM common.jl:224-241
    const double* prm = a.r_prm;
    double ktt;
    if (a.r_ktype == ISO_SE) ktt = prm[PRM_V];
    else if (a.r_ktype == ARD_SE) ktt = prm[PRM_V] * (double)a.D;
    else {
      ktt = 0.0;
      for (int d = 0; d < a.D; d++) { const double xv = a.xtest[(int64_t)d * a.T + p]; ktt += (a.r_ktype == ISO_LINEAR ? prm[PRM_COEF] : prm[PRM_COEF + d]) * xv * xv; }
    }
    s_prior = ktt + prm[PRM_ETA];
    C = 1.0 / s_prior;
  }
  const double beta_g = 1.0 / (double)RK;
  int r = 0;
  for (int kr = 0; kr < (a.mode == 1 ? 1 : RK); kr++) {
    // _predictPoE of the root (PoE) or of one child of the root (gPoE / rBCM), iteratively
    int fnode[MIX_FRAMES], fk[MIX_FRAMES];
    double fm[MIX_FRAMES], ft[MIX_FRAMES];
    int nf = 0, node = (a.mode == 1) ? root : t.child_idx[rc0 + kr];
    double m_ = 0, t_ = 0;
    for (;;) {
      while (t.type[node] != 0) { fnode[nf] = node; fk[nf] = 0; fm[nf] = 0.0; ft[nf] = 0.0; nf++; node = t.child_idx[t.child_ptr[node]]; }
      { const int o = rc[r++]; m_ = a.lmu[o]; t_ = 1.0 / a.lvar[o]; }            // common.jl:145-149
      bool done = false;
      for (;;) {
        if (nf == 0) { done = true; break; }
        const int f = nf - 1, nd = fnode[f], c0 = t.child_ptr[nd], K = t.child_ptr[nd + 1] - c0;
        ft[f] += t_; fm[f] += t_ * m_;
        fk[f]++;
        if (fk[f] < K) { node = t.child_idx[c0 + fk[f]]; break; }
        m_ = fm[f] / ft[f]; t_ = ft[f];
        nf--;
      }
      if (done) break;
    }
    if (a.mode == 1) { a.mu[p] = m_; a.var[p] = 1.0 / t_; return; }
    if (a.mode == 2) { Tt += beta_g * t_; M += beta_g * t_ * m_; }                 // common.jl:211-222
    else {
      const double s_ = 1.0 / t_;
      const double beta = 0.5 * (log(s_prior) - log(s_));
      C = C + (beta * t_) - (beta / s_prior);
      M = M + m_ * (beta * t_);
    }
  }
  if (a.mode == 2) { a.mu[p] = M / Tt; a.var[p] = 1.0 / Tt; }
  else { a.mu[p] = M / C; a.var[p] = 1.0 / C; }
}

}  // namespace dsm
