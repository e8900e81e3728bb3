// Argument blocks of the device-side tree passes and the fused optimiser step (tree.cuh), host-visible.
#pragma once
#include "common.cuh"
#include "route_args.h"

namespace dsm {

// mll! (optimize.jl:27-39) and the down-pass nabla-mll! (:42-150) over the flattened region graph, ONE CTA.
struct TreeEvalArgs {
  DevTree t;
  const int* up_ptr; const int* up_nodes; int n_up;        // nodes grouped by height (leaves first)
  const int* dn_ptr; const int* dn_nodes; int n_dn;        // nodes grouped by depth (root first)
  const int* leaf_dfs;                                     // leaves in getLeaves order (the reference's summation order)
  const int* leaf_node;                                    // node id of every leaf
  const int* leaf_goff;                                    // first gradient component of every leaf (kernel-mixture slices)
  const int* leaf_np;                                      // nparams of every leaf's kernel
  int L, H;
  const double* rows; int row_width;                       // per-leaf rows [mll(gp), nabla-mll(gp)...]
  const double* leaf_scale;                                // finetune weights D[g,:] or null
  double* ell; double* dpar; double* lrho; double* w;      // scratch: [n_nodes] x 3, [L]
  double* out;                                             // [1 + H]: mll(root), gradient
};

// theta -> derived per-slot parameter blocks (what dsmgp_set_params does on the host): one thread per slot
struct DeriveArgs {
  const LeafMeta* meta; int nslots;
  const double* theta;                                     // [H] global theta (concatenated per kernel)
  const int* koff;                                         // theta offset per kernel id
  double* prm; int pstride;
};

// Flux.Optimise.apply! + `hyp += grad` (optimisers.jl:78-79): Descent / ADAM / RMSProp, gradient ASCENT
struct OptArgs {
  int optimiser; double eta, beta1, beta2; int state_by_identity;
  int H;
  double* theta;                                           // [H] in / out
  const double* out;                                       // [1 + H] mll, gradient of this iteration
  double* mt; double* vt; double* acc; double* bp;         // optimiser state ([H] x 3, [2] running beta powers)
  double* ell; double* hist;                               // [iterations] LML trace, [iterations][H] theta used by each iteration
  int* it;                                                 // device iteration counter
};

void launch_tree_eval(const TreeEvalArgs& a, cudaStream_t st);
void launch_derive(const DeriveArgs& a, cudaStream_t st);
void launch_opt_step(const OptArgs& a, cudaStream_t st);

}  // namespace dsm
