// Argument blocks of the device-side routing / mixing kernels (route.cuh), host-visible.
#pragma once
#include "common.cuh"

namespace dsm {

constexpr int ROUTE_STACK = 96;     // pending nodes of the depth-first walk (checked on the host: depth * max fan-out)
constexpr int MIX_KMAX = 8;         // children per sum node the device mixer supports (else: host mixer)
constexpr int MIX_FRAMES = 8;       // nested sum nodes (DSMGP) / split nodes (PoE) on one root-to-leaf path

struct DevTree {
  int n_nodes, root;
  const int* type;            // NODE_LEAF 0, SPLIT 1, SUM 2, KSUM 3
  const int* child_ptr;       // [n_nodes + 1]
  const int* child_idx;
  const int* leaf_of_node;    // global leaf or -1
  const int* split_dim;
  const int* split_ptr;       // [n_nodes + 1]
  const double* split_val;
};

struct RouteArgs {
  DevTree t;
  const double* xtest; int64_t T; int D;     // T x D column-major
  int poe;                                   // 1: split nodes send a point to every child (PoE family)
  int R;                                     // experts a point can reach at most (row length of `reach`)
  int* cnt;                                  // [L] points per expert (count pass)
  const int* ooff;                           // [L] first output position of every expert (fill pass), padded to BLK
  int* fill;                                 // [L] running position (fill pass)
  int* pidx;                                 // [total padded] test point of every (expert, position), -1 on the padding
  int* reach;                                // [T][R] output position of the r-th expert of a point, -1 unused
  int* err;                                  // 1: non-finite input, 2: point outside every split interval
};

struct MixArgs {
  DevTree t;
  const double* xtest; int64_t T; int D;
  int mode;                      // DSMGP 0, PoE 1, gPoE 2, rBCM 3
  int R;
  const int* reach;              // [T][R]
  const double* lmu;             // per (expert, position): prediction(gp, x) mean and variance (gaussianprocess.jl:110-137)
  const double* lvar;
  const double* logw;            // sum-node log-weights, CSR by child_ptr (update! / infer! / reset_weights!)
  // rBCM prior: the left-most expert's kernel (common.jl:226-227)
  int r_ktype; const double* r_prm;
  double* mu; double* var;       // [T]
};

}  // namespace dsm
