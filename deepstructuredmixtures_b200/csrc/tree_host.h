// Host-side passes over the flattened region graph (O(#nodes) scalar work; no GPU involved).
//
//   up-pass     mll!(node, l)            optimize.jl:27-39
//   down-pass   nabla-mll!(node, ...)    optimize.jl:42-89 and the finetune variant :92-150
//   weights     update!(node)            common.jl:323-334
//   routing     getchild                 common.jl:101-122
//   mixing      _minpredict / _predict / predict / predictPoE / predictgPoE / predictrBCM   common.jl:134-307
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <string>
#include <vector>
#include "../../include/dsmgp.h"

namespace dsm {

struct HostTree {
  int64_t n_nodes = 0, root = 0;
  std::vector<int32_t> type;
  std::vector<int64_t> child_ptr, child_idx, leaf_of_node, split_ptr;
  std::vector<int32_t> split_dim;
  std::vector<double> split_val;

  bool load(const dsmgp_tree* t, int64_t L, std::string& err) {
    if (!t || t->n_nodes <= 0) { err = "tree: n_nodes <= 0"; return false; }
    n_nodes = t->n_nodes; root = t->root;
    if (root < 0 || root >= n_nodes) { err = "tree: root out of range"; return false; }
    type.assign(t->node_type, t->node_type + n_nodes);
    child_ptr.assign(t->child_ptr, t->child_ptr + n_nodes + 1);
    child_idx.assign(t->child_idx, t->child_idx + child_ptr[n_nodes]);
    leaf_of_node.assign(t->leaf_of_node, t->leaf_of_node + n_nodes);
    split_dim.assign(t->split_dim, t->split_dim + n_nodes);
    split_ptr.assign(t->split_ptr, t->split_ptr + n_nodes + 1);
    split_val.assign(t->split_val, t->split_val + split_ptr[n_nodes]);
    for (int64_t i = 0; i < n_nodes; i++) {
      if (type[i] < 0 || type[i] > 3) { err = "tree: bad node type"; return false; }
      for (int64_t c = child_ptr[i]; c < child_ptr[i + 1]; c++)
        if (child_idx[c] < 0 || child_idx[c] >= i) { err = "tree: children must precede parents"; return false; }
      if (type[i] == DSMGP_NODE_LEAF) {
        if (leaf_of_node[i] < 0 || leaf_of_node[i] >= L) { err = "tree: leaf_of_node out of range"; return false; }
      } else if (child_ptr[i + 1] == child_ptr[i]) { err = "tree: inner node without children"; return false; }
      if (type[i] == DSMGP_NODE_SPLIT && split_ptr[i + 1] - split_ptr[i] != child_ptr[i + 1] - child_ptr[i]) {
        err = "tree: split node needs one threshold per child"; return false;
      }
    }
    return true;
  }
  int64_t nchild(int64_t i) const { return child_ptr[i + 1] - child_ptr[i]; }
  int64_t child(int64_t i, int64_t k) const { return child_idx[child_ptr[i] + k]; }
};

inline double logsumexp(const double* v, int64_t n) {   // StatsFuns.logsumexp
  double m = -std::numeric_limits<double>::infinity();
  for (int64_t i = 0; i < n; i++) if (v[i] > m || std::isnan(v[i])) m = v[i];
  if (!std::isfinite(m)) return m;
  double s = 0.0;
  for (int64_t i = 0; i < n; i++) s += std::exp(v[i] - m);
  return m + std::log(s);
}

// optimize.jl:27-39.  rows: per-leaf rows (row_width doubles each), element 0 = mll(gp).
inline void up_pass(const HostTree& t, const double* rows, int64_t rw, double* ell) {
  std::vector<double> tmp;
  for (int64_t i = 0; i < t.n_nodes; i++) {
    const int ty = t.type[i];
    if (ty == DSMGP_NODE_LEAF) ell[i] = rows[t.leaf_of_node[i] * rw];
    else if (ty == DSMGP_NODE_SPLIT) {
      double s = ell[t.child(i, 0)];
      for (int64_t k = 1; k < t.nchild(i); k++) s = s + ell[t.child(i, k)];
      ell[i] = s;
    } else {
      const int64_t K = t.nchild(i);
      tmp.resize(K);
      for (int64_t k = 0; k < K; k++) tmp[k] = -std::log((double)K) + ell[t.child(i, k)];
      ell[i] = logsumexp(tmp.data(), K);
    }
  }
}

struct DownCtx {
  const HostTree* t;
  const double* rows; int64_t rw;
  const double* ell; double logS;
  const double* leaf_scale;             // finetune D[g,:] or null
  const int32_t* leaf_kid; const int64_t* koff; const int32_t* knp;   // per kernel: theta offset, nparams
  double* grad;
};
// optimize.jl:42-89 / :92-150
inline void down_pass(const DownCtx& c, int64_t node, double dparent, double lrho, int64_t goff) {
  const HostTree& t = *c.t;
  const int ty = t.type[node];
  if (ty == DSMGP_NODE_LEAF) {
    const int64_t l = t.leaf_of_node[node];
    double w = std::exp(-c.logS + lrho + c.ell[node] + dparent);
    if (c.leaf_scale) w = w * c.leaf_scale[l];
    const int np = c.knp[c.leaf_kid[l]];
    const double* g = c.rows + l * c.rw + 1;
    for (int h = 0; h < np; h++) c.grad[goff + h] += g[h] * w;
  } else if (ty == DSMGP_NODE_SPLIT) {
    for (int64_t k = 0; k < t.nchild(node); k++) {
      const int64_t ch = t.child(node, k);
      const double lp = c.ell[node] - c.ell[ch];
      down_pass(c, ch, dparent + lp, lrho, goff);
    }
  } else if (ty == DSMGP_NODE_SUM) {
    const double lK = std::log((double)t.nchild(node));
    for (int64_t k = 0; k < t.nchild(node); k++) down_pass(c, t.child(node, k), -lK + dparent, lK + lrho, goff);
  } else {   // kernel mixture: slices of grad per kernel (optimize.jl:76-89)
    int64_t off = 0;
    for (int64_t k = 0; k < t.nchild(node); k++) {
      const int64_t ch = t.child(node, k);
      down_pass(c, ch, dparent, lrho, goff + off);
      off += c.knp[c.leaf_kid[t.leaf_of_node[ch]]];
    }
  }
}

// common.jl:323-334.  logw: CSR by child_ptr over ALL nodes (entries of non-sum nodes untouched).
inline double update_weights(const HostTree& t, const double* rows, int64_t rw, double* logw, double* zval) {
  std::vector<double> val(t.n_nodes);
  for (int64_t i = 0; i < t.n_nodes; i++) {
    const int ty = t.type[i];
    if (ty == DSMGP_NODE_LEAF) val[i] = rows[t.leaf_of_node[i] * rw];
    else if (ty == DSMGP_NODE_SPLIT) {
      double s = val[t.child(i, 0)];
      for (int64_t k = 1; k < t.nchild(i); k++) s = s + val[t.child(i, k)];
      val[i] = s;
    } else {
      const int64_t K = t.nchild(i);
      double* lw = logw + t.child_ptr[i];
      for (int64_t k = 0; k < K; k++) lw[k] = -std::log((double)K) + val[t.child(i, k)];
      const double z = logsumexp(lw, K);
      for (int64_t k = 0; k < K; k++) lw[k] = lw[k] - z;
      val[i] = z;
    }
  }
  if (zval) *zval = val[t.root];
  return val[t.root];
}

// infer! common.jl:336-355: kernel-mixture sum nodes (children are GPNodes, :339-345) keep normalised posterior weights, sum
// nodes over sub-trees are reset to -log K after their evidence z is computed (:347-353).
inline double infer_weights(const HostTree& t, const double* rows, int64_t rw, double* logw, double* zval) {
  std::vector<double> val(t.n_nodes);
  for (int64_t i = 0; i < t.n_nodes; i++) {
    const int ty = t.type[i];
    if (ty == DSMGP_NODE_LEAF) val[i] = rows[t.leaf_of_node[i] * rw];
    else if (ty == DSMGP_NODE_SPLIT) {
      double s = val[t.child(i, 0)];
      for (int64_t k = 1; k < t.nchild(i); k++) s = s + val[t.child(i, k)];
      val[i] = s;
    } else {
      const int64_t K = t.nchild(i);
      double* lw = logw + t.child_ptr[i];
      for (int64_t k = 0; k < K; k++) lw[k] = -std::log((double)K) + val[t.child(i, k)];
      const double z = logsumexp(lw, K);
      if (ty == DSMGP_NODE_KSUM) { for (int64_t k = 0; k < K; k++) lw[k] = lw[k] - z; }
      else { for (int64_t k = 0; k < K; k++) lw[k] = -std::log((double)K); }
      val[i] = z;
    }
  }
  if (zval) *zval = val[t.root];
  return val[t.root];
}

// reset_weights! common.jl:357-363
inline void reset_weights(const HostTree& t, double* logw) {
  for (int64_t i = 0; i < t.n_nodes; i++)
    if (t.type[i] >= DSMGP_NODE_SUM) for (int64_t k = 0; k < t.nchild(i); k++) logw[t.child_ptr[i] + k] = -std::log((double)t.nchild(i));
}

// common.jl:101-122 for one point; returns child position or -1 (x above the last threshold / NaN)
inline int64_t getchild(const HostTree& t, int64_t node, const double* xtest, int64_t T, int64_t p) {
  const int d = t.split_dim[node];
  const double xv = xtest[(int64_t)d * T + p];
  const int64_t K = t.nchild(node);
  const double* s = t.split_val.data() + t.split_ptr[node];
  for (int64_t k = 0; k < K; k++) {
    const bool ok = (k == 0) ? (xv <= s[0]) : ((xv <= s[k]) && (xv > s[k - 1]));
    if (ok) return k;
  }
  return -1;
}

}  // namespace dsm
