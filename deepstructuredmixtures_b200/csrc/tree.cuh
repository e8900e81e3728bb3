// Tree passes and the optimiser step on the device: a whole train! iteration (optimisers.jl:43-79) without a host round trip.
//
//   tree_eval_kernel   mll!(spn, l) optimize.jl:27-39 (leaf: row[0]; split: sum in child order; sum: logsumexp(-log K + child))
//                      nabla-mll!(spn, 0, 0, l, l[root], grad) optimize.jl:42-89 (+ the finetune weights :92-150):
//                      leaf weight w = exp(-logS + lrho + l_leaf + dparent), grad += w * nabla-mll(leaf) in getLeaves order
//   derive_kernel      setparams! -> the derived parameter blocks the kernels read (kernels.jl:68-73, gaussianprocess.jl:39)
//   opt_step_kernel    Flux.Optimise.apply!(optim, hyp, grad); hyp += grad   (optimisers.jl:78-79; Descent / ADAM / RMSProp)
// The graph is O(#nodes) scalar work: one CTA walks it level by level; the gradient is summed by one thread per component in
// the reference's own (depth-first leaf) order, so the result does not depend on the launch configuration.
#pragma once
#include "common.cuh"
#include "tree_args.h"

namespace dsm {

__global__ void __launch_bounds__(256) tree_eval_kernel(TreeEvalArgs a) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const DevTree& t = a.t;
  // up-pass, level by level (children before parents)
  for (int lv = 0; lv < a.n_up; lv++) {
    for (int q = a.up_ptr[lv] + tid; q < a.up_ptr[lv + 1]; q += nt) {
      const int node = a.up_nodes[q], ty = t.type[node];
      const int c0 = t.child_ptr[node], K = t.child_ptr[node + 1] - c0;
      double v;
      if (ty == 0) v = a.rows[(int64_t)t.leaf_of_node[node] * a.row_width];
      else if (ty == 1) {
        v = a.ell[t.child_idx[c0]];
        for (int k = 1; k < K; k++) v = v + a.ell[t.child_idx[c0 + k]];
      } else {                                            // StatsFuns.logsumexp(-log K + child)
        const double lK = log((double)K);
        double m = -INFINITY;
        bool nan_seen = false;
        for (int k = 0; k < K; k++) { const double x = -lK + a.ell[t.child_idx[c0 + k]]; if (isnan(x)) nan_seen = true; m = fmax(m, x); }
        if (nan_seen) v = NAN;
        else if (isinf(m)) v = m;
        else {
          double s = 0.0;
          for (int k = 0; k < K; k++) s += exp((-lK + a.ell[t.child_idx[c0 + k]]) - m);
          v = m + log(s);
        }
      }
      a.ell[node] = v;
    }
    __syncthreads();
  }
  const double logS = a.ell[t.root];
  if (tid == 0) { a.dpar[t.root] = 0.0; a.lrho[t.root] = 0.0; }
  __syncthreads();
  // down-pass, level by level (parents before children)
  for (int lv = 0; lv < a.n_dn; lv++) {
    for (int q = a.dn_ptr[lv] + tid; q < a.dn_ptr[lv + 1]; q += nt) {
      const int node = a.dn_nodes[q], ty = t.type[node];
      if (ty == 0) continue;
      const int c0 = t.child_ptr[node], K = t.child_ptr[node + 1] - c0;
      const double dp = a.dpar[node], lr = a.lrho[node];
      const double lK = log((double)K);
      for (int k = 0; k < K; k++) {
        const int ch = t.child_idx[c0 + k];
        if (ty == 1) { a.dpar[ch] = dp + (a.ell[node] - a.ell[ch]); a.lrho[ch] = lr; }          // optimize.jl:58-61
        else if (ty == 2) { a.dpar[ch] = -lK + dp; a.lrho[ch] = lK + lr; }                      // :70-73
        else { a.dpar[ch] = dp; a.lrho[ch] = lr; }                                              // kernel mixture :76-89
      }
    }
    __syncthreads();
  }
  for (int l = tid; l < a.L; l += nt) {
    const int node = a.leaf_node[l];
    double w = exp(-logS + a.lrho[node] + a.ell[node] + a.dpar[node]);                          // :48
    if (a.leaf_scale != nullptr) w = w * a.leaf_scale[l];                                       // :101
    a.w[l] = w;
  }
  __syncthreads();
  for (int c = tid; c < a.H; c += nt) {
    double g = 0.0;
    for (int q = 0; q < a.L; q++) {
      const int l = a.leaf_dfs[q];
      const int j = c - a.leaf_goff[l];
      if (j >= 0 && j < a.leaf_np[l]) g += a.rows[(int64_t)l * a.row_width + 1 + j] * a.w[l];
    }
    a.out[1 + c] = g;
  }
  if (tid == 0) a.out[0] = logS;
}

__global__ void derive_kernel(DeriveArgs a) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.nslots) return;
  const LeafMeta m = a.meta[s];
  const double* th = a.theta + a.koff[m.kid];
  double* prm = a.prm + (int64_t)s * a.pstride;
  const bool se = (m.ktype == ISO_SE || m.ktype == ARD_SE);
  const double logs = th[m.nl], logn = th[m.nl + 1];
  prm[PRM_V] = se ? exp(2.0 * logs) : 1.0;
  prm[PRM_S] = se ? exp(logs) : 1.0;
  const double eta = exp(2.0 * logn);
  prm[PRM_ETA] = eta;
  prm[PRM_C] = eta + 1e-8;
  for (int d = 0; d < m.nl; d++) {
    const double l = exp(th[d]);
    const double l2 = l * l;
    prm[PRM_COEF + d] = se ? -0.5 / l2 : 1.0 / l2;
  }
}

__global__ void opt_step_kernel(OptArgs a) {
  const int k = threadIdx.x;
  const int it = *a.it;
  if (k < a.H) a.hist[(int64_t)it * a.H + k] = a.theta[k];
  if (k == 0) a.ell[it] = a.out[0];
  if (k < a.H) {
    double d = a.out[1 + k];
    double bp1 = a.bp[0], bp2 = a.bp[1];
    double mt = a.mt[k], vt = a.vt[k], acc = a.acc[k];
    if (a.state_by_identity) { mt = 0.0; vt = 0.0; acc = 0.0; bp1 = a.beta1; bp2 = a.beta2; }   // a fresh state every iteration (App. B Q9)
    if (a.optimiser == 0) d *= a.eta;
    else if (a.optimiser == 1) {
      mt = a.beta1 * mt + (1.0 - a.beta1) * d;
      vt = a.beta2 * vt + (1.0 - a.beta2) * d * d;
      d = mt / (1.0 - bp1) / (sqrt(vt / (1.0 - bp2)) + 1e-8) * a.eta;
    } else {
      acc = a.beta1 * acc + (1.0 - a.beta1) * d * d;
      d = d * (a.eta / (sqrt(acc) + 1e-8));
    }
    a.mt[k] = mt; a.vt[k] = vt; a.acc[k] = acc;
    a.theta[k] = a.theta[k] + d;                                                                // hyp += grad (ASCENT)
  }
  __syncthreads();
  if (k == 0) {
    if (a.optimiser == 1 && !a.state_by_identity) { a.bp[0] *= a.beta1; a.bp[1] *= a.beta2; }
    *a.it = it + 1;
  }
}

}  // namespace dsm
