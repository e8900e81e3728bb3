// libdsmgp.so : the train! loop inside the library (optimisers.jl:40-83), SURVEY 8f rank 4.
//
// Two implementations with identical semantics:
//   * host loop      every iteration = dsmgp_eval (one stream synchronisation + read-back of the row table + host tree passes)
//                    followed by the Flux step on the host;
//   * fused loop     theta, the optimiser state and the LML trace live on the device; ONE CUDA graph holds a whole iteration
//                    (theta -> derived parameters -> Gram -> Cholesky -> inverse -> rows -> tree passes -> Flux step) and is
//                    launched back to back.  The host only looks at the trace every CHUNK iterations to apply the early-stopping
//                    rule (optimisers.jl:53-66); theta of every iteration is kept, so stopping "before the update of iteration
//                    t" is exact even when the device has already run past t.  For launch-bound models (README example:
//                    66 experts of 5-20 points) the per-iteration cost drops from the host round trip to the kernels' own time.
#include "handle.h"
#include "tree_args.h"

using namespace dsm;
#define g_create_error (dsm::create_error())

// train!(spn, D, gpmap, optim; iterations, lambda, earlystop) optimisers.jl:40-83 as ONE call: the loop stays inside the
// library (no per-iteration host round trip through the binding).  Optimisers = Flux.Optimise Descent / ADAM / RMSProp;
// `state_by_identity` reproduces the reference's `hyp += grad` rebinding, which gives apply! a fresh state every iteration
// (SURVEY App. B Q9).  The update is gradient ASCENT.  Returns the number of iterations executed in *n_done.
static int32_t train_host_loop(dsmgp_handle* h, int32_t optimiser, double eta, double beta1, double beta2,
                               int32_t state_by_identity, int64_t iterations, double lambda, int64_t earlystop,
                               double* theta, double* ell, int64_t* n_done) {
  if (!h) return DSMGP_ERR_ARG;
  if (!theta || !ell || iterations <= 0 || optimiser < 0 || optimiser > 2) { h->err = "train: bad argument"; return DSMGP_ERR_ARG; }
  if (h->opts.world != 1) { h->err = "train: single-process handles only"; return DSMGP_ERR_STATE; }
  const int64_t H = h->H;
  std::vector<double> hyp(theta, theta + H), grad(H), mt(H, 0.0), vt(H, 0.0), acc(H, 0.0);
  double bp1 = beta1, bp2 = beta2;
  int64_t c = 0, it = 0;
  if (n_done) *n_done = 0;
  for (it = 0; it < iterations; it++) {
    double lml = 0.0;
    int32_t rc = dsmgp_eval(h, hyp.data(), H, nullptr, &lml, grad.data(), nullptr);     // :43-49, 68-77
    if (rc) return rc;
    ell[it] = lml;
    double delta = std::numeric_limits<double>::infinity();
    if (it >= 10) { double mean = 0.0; for (int64_t k = it - 9; k < it; k++) mean += ell[k]; delta = std::fabs(ell[it] - mean / 9.0); }   // :53
    c = (delta < lambda) ? c + 1 : 0;                                                    // :57-61
    if (c >= earlystop) { it++; break; }                                                 // :63-66 (returns before the update)
    if (state_by_identity) { std::fill(mt.begin(), mt.end(), 0.0); std::fill(vt.begin(), vt.end(), 0.0); std::fill(acc.begin(), acc.end(), 0.0); bp1 = beta1; bp2 = beta2; }
    for (int64_t k = 0; k < H; k++) {                                                    // Flux.Optimise.apply!  :78
      double d = grad[k];
      if (optimiser == 0) d *= eta;
      else if (optimiser == 1) {
        mt[k] = beta1 * mt[k] + (1.0 - beta1) * d;
        vt[k] = beta2 * vt[k] + (1.0 - beta2) * d * d;
        d = mt[k] / (1.0 - bp1) / (std::sqrt(vt[k] / (1.0 - bp2)) + 1e-8) * eta;
      } else {
        acc[k] = beta1 * acc[k] + (1.0 - beta1) * d * d;                                 // RMSProp: beta1 = rho
        d = d * (eta / (std::sqrt(acc[k]) + 1e-8));
      }
      hyp[k] = hyp[k] + d;                                                               // :79
    }
    if (optimiser == 1) { bp1 *= beta1; bp2 *= beta2; }
  }
  std::copy(hyp.begin(), hyp.end(), theta);
  if (n_done) *n_done = it;
  if (it >= iterations) {                                                                // :82-83 final setparams! + fit!
    int32_t rc = dsmgp_set_params(h, hyp.data(), H);
    if (rc) return rc;
    if ((rc = run_pipeline(h, false))) return rc;
    return fetch_rows(h);
  }
  return DSMGP_OK;
}


// level lists of the flattened tree for tree_eval_kernel (uploaded once)
static int32_t ensure_levels(dsmgp_handle* h) {
  if (h->lvl_on_device) return DSMGP_OK;
  const HostTree& t = h->tree;
  const int64_t nn = t.n_nodes, L = h->L;
  std::vector<int> height(nn, 0), depth(nn, 0);
  for (int64_t i = 0; i < nn; i++) for (int64_t k = 0; k < t.nchild(i); k++) height[i] = std::max(height[i], height[t.child(i, k)] + 1);
  for (int64_t i = nn - 1; i >= 0; i--) for (int64_t k = 0; k < t.nchild(i); k++) depth[t.child(i, k)] = depth[i] + 1;   // parents have larger ids
  // nodes that do not hang below the root (none in practice) keep depth 0 and are harmless
  const int nu = *std::max_element(height.begin(), height.end()) + 1, nd = *std::max_element(depth.begin(), depth.end()) + 1;
  std::vector<int> up_ptr(nu + 1, 0), dn_ptr(nd + 1, 0), up_nodes(nn), dn_nodes(nn);
  for (int64_t i = 0; i < nn; i++) { up_ptr[height[i] + 1]++; dn_ptr[depth[i] + 1]++; }
  for (int i = 0; i < nu; i++) up_ptr[i + 1] += up_ptr[i];
  for (int i = 0; i < nd; i++) dn_ptr[i + 1] += dn_ptr[i];
  { std::vector<int> f(up_ptr.begin(), up_ptr.end() - 1), g(dn_ptr.begin(), dn_ptr.end() - 1);
    for (int64_t i = 0; i < nn; i++) { up_nodes[f[height[i]]++] = (int)i; dn_nodes[g[depth[i]]++] = (int)i; } }
  // leaves in getLeaves (depth-first, child order) order; gradient slice of every leaf (kernel-mixture sums slice by kernel)
  std::vector<int> leaf_dfs, leaf_node(L, 0), leaf_goff(L, 0), leaf_np(L, 0);
  struct Rec { const HostTree& t; dsmgp_handle* h; std::vector<int>& dfs; std::vector<int>& goff;
    void run(int64_t node, int off) {
      const int ty = t.type[node];
      if (ty == DSMGP_NODE_LEAF) { dfs.push_back((int)t.leaf_of_node[node]); goff[t.leaf_of_node[node]] = off; return; }
      int o = 0;
      for (int64_t k = 0; k < t.nchild(node); k++) {
        const int64_t ch = t.child(node, k);
        run(ch, ty == DSMGP_NODE_KSUM ? off + o : off);
        if (ty == DSMGP_NODE_KSUM) o += h->knp[h->leaf_kid[t.leaf_of_node[ch]]];
      }
    } } rec{t, h, leaf_dfs, leaf_goff};
  rec.run(t.root, 0);
  for (int64_t i = 0; i < nn; i++) if (t.type[i] == DSMGP_NODE_LEAF) leaf_node[t.leaf_of_node[i]] = (int)i;
  for (int64_t l = 0; l < L; l++) leaf_np[l] = h->knp[h->leaf_kid[l]];
  std::vector<int> koff(h->nk);
  for (int k = 0; k < h->nk; k++) koff[k] = (int)h->koff[k];
  std::vector<int> buf;
  auto put = [&](int slot, const std::vector<int>& v) { h->lvl_off[slot] = buf.size(); buf.insert(buf.end(), v.begin(), v.end()); };
  put(0, up_ptr); put(1, up_nodes); put(2, dn_ptr); put(3, dn_nodes); put(4, leaf_dfs); put(5, leaf_node); put(6, leaf_goff); put(7, leaf_np);
  const size_t koff_at = buf.size();
  buf.insert(buf.end(), koff.begin(), koff.end());
  CUDA_TRY(h, h->t_lvl.alloc(buf.size() + 1));
  CUDA_TRY(h, cudaMemcpy(h->t_lvl.p, buf.data(), buf.size() * sizeof(int), cudaMemcpyHostToDevice));
  h->n_up = nu; h->n_dn = nd;
  h->lvl_on_device = true;
  (void)koff_at;
  return DSMGP_OK;
}

static int32_t train_fused(dsmgp_handle* h, int32_t optimiser, double eta, double beta1, double beta2, int32_t state_by_identity,
                           int64_t iterations, double lambda, int64_t earlystop, double* theta, double* ell, int64_t* n_done) {
  cudaStream_t st = h->stream;
  const int64_t H = h->H, L = h->L, nn = h->tree.n_nodes;
  int32_t rc;
  if ((rc = ensure_dev_tree(h))) return rc;
  if ((rc = ensure_levels(h))) return rc;
  const int ns = (int)h->slot_leaf.size();
  // device state: theta H | out 1+H | mt H | vt H | acc H | bp 2 | ell nn | dpar nn | lrho nn | w L | trace iters | hist iters*H
  const size_t o_theta = 0, o_out = o_theta + H, o_mt = o_out + 1 + H, o_vt = o_mt + H, o_acc = o_vt + H, o_bp = o_acc + H,
               o_ell = o_bp + 2, o_dpar = o_ell + nn, o_lrho = o_dpar + nn, o_w = o_lrho + nn, o_tr = o_w + L,
               o_hist = o_tr + iterations, total = o_hist + (size_t)iterations * H;
  CUDA_TRY(h, h->g_dbl.ensure(total));
  CUDA_TRY(h, h->g_int.ensure(4));
  double* g = h->g_dbl.p;
  CUDA_TRY(h, cudaMemsetAsync(g, 0, total * sizeof(double), st));
  CUDA_TRY(h, cudaMemsetAsync(h->g_int.p, 0, 4 * sizeof(int), st));
  CUDA_TRY(h, cudaMemcpyAsync(g + o_theta, theta, H * sizeof(double), cudaMemcpyHostToDevice, st));
  const double bp0[2] = {beta1, beta2};
  CUDA_TRY(h, cudaMemcpyAsync(g + o_bp, bp0, 2 * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaMemsetAsync(h->d_counter.p + GERR, 0, sizeof(int), st));
  CUDA_TRY(h, cudaStreamSynchronize(st));
  const int* lv = h->t_lvl.p;
  const size_t koff_at = h->lvl_off[7] + L;
  DeriveArgs da{h->d_meta.p, ns, g + o_theta, lv + koff_at, h->d_prm.p, h->pstride};
  TreeEvalArgs ta{dev_tree(h), lv + h->lvl_off[0], lv + h->lvl_off[1], h->n_up, lv + h->lvl_off[2], lv + h->lvl_off[3], h->n_dn,
                  lv + h->lvl_off[4], lv + h->lvl_off[5], lv + h->lvl_off[6], lv + h->lvl_off[7], (int)L, (int)H,
                  h->d_rows.p, h->row_width, nullptr, g + o_ell, g + o_dpar, g + o_lrho, g + o_w, g + o_out};
  OptArgs oa{optimiser, eta, beta1, beta2, state_by_identity, (int)H, g + o_theta, g + o_out, g + o_mt, g + o_vt, g + o_acc, g + o_bp,
             g + o_tr, g + o_hist, h->g_int.p};
  // ---- one iteration as a CUDA graph
  cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
  CUDA_TRY(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  h->capturing = true;
  launch_derive(da, st);
  rc = run_pipeline(h, true, nullptr, /*defer_sync=*/true, /*naive=*/true, /*first=*/false);
  launch_tree_eval(ta, st);
  launch_opt_step(oa, st);
  h->capturing = false;
  cudaError_t ce = cudaStreamEndCapture(st, &graph);
  if (rc != DSMGP_OK || ce != cudaSuccess) {
    if (graph) cudaGraphDestroy(graph);
    if (rc == DSMGP_OK) { h->err = std::string("train: graph capture failed: ") + cudaGetErrorString(ce); rc = DSMGP_ERR_CUDA; }
    cudaGetLastError();
    return rc;
  }
  ce = cudaGraphInstantiate(&exec, graph, 0);
  if (ce != cudaSuccess) { cudaGraphDestroy(graph); h->err = std::string("train: cudaGraphInstantiate: ") + cudaGetErrorString(ce); return DSMGP_ERR_CUDA; }
  // ---- launch back to back; the host applies the early-stopping rule on the trace every CHUNK iterations
  const int64_t CHUNK = 32;
  int64_t launched = 0, checked = 0, c = 0, stop_at = -1;
  rc = DSMGP_OK;
  while (launched < iterations && stop_at < 0) {
    const int64_t n = std::min<int64_t>(CHUNK, iterations - launched);
    for (int64_t i = 0; i < n; i++) if ((ce = cudaGraphLaunch(exec, st)) != cudaSuccess) break;
    if (ce != cudaSuccess) { h->err = std::string("train: cudaGraphLaunch: ") + cudaGetErrorString(ce); rc = DSMGP_ERR_CUDA; break; }
    launched += n;
    int gerr = 0;
    cudaMemcpyAsync(ell + checked, g + o_tr + checked, (launched - checked) * sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&gerr, h->d_counter.p + GERR, sizeof(int), cudaMemcpyDeviceToHost, st);
    if ((ce = cudaStreamSynchronize(st)) != cudaSuccess) { h->err = std::string("train: ") + cudaGetErrorString(ce); rc = DSMGP_ERR_CUDA; break; }
    if (gerr != 0) { h->err = "device scheduler timeout (code " + std::to_string(gerr) + ")"; rc = DSMGP_ERR_STATE; break; }
    for (; checked < launched; checked++) {                                                // optimisers.jl:53-66
      const int64_t it = checked;
      double delta = std::numeric_limits<double>::infinity();
      if (it >= 10) { double mean = 0.0; for (int64_t k = it - 9; k < it; k++) mean += ell[k]; delta = std::fabs(ell[it] - mean / 9.0); }
      c = (delta < lambda) ? c + 1 : 0;
      if (c >= earlystop) { stop_at = it; break; }
    }
  }
  cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
  if (rc) return rc;
  std::vector<double> hyp(H);
  if (stop_at >= 0) {
    // returned BEFORE the update of iteration stop_at: theta of that iteration, handle left evaluated at it
    CUDA_TRY(h, cudaMemcpy(hyp.data(), g + o_hist + (size_t)stop_at * H, H * sizeof(double), cudaMemcpyDeviceToHost));
    std::copy(hyp.begin(), hyp.end(), theta);
    if (n_done) *n_done = stop_at + 1;
    if ((rc = dsmgp_set_params(h, hyp.data(), H))) return rc;
    if ((rc = run_pipeline(h, true))) return rc;
    return fetch_rows(h);
  }
  CUDA_TRY(h, cudaMemcpy(hyp.data(), g + o_theta, H * sizeof(double), cudaMemcpyDeviceToHost));
  std::copy(hyp.begin(), hyp.end(), theta);
  if (n_done) *n_done = iterations;
  if ((rc = dsmgp_set_params(h, hyp.data(), H))) return rc;                                // :82-83 final setparams! + fit!
  if ((rc = run_pipeline(h, false))) return rc;
  return fetch_rows(h);
}

extern "C" int32_t dsmgp_train(dsmgp_handle* h, int32_t optimiser, double eta, double beta1, double beta2,
                               int32_t state_by_identity, int64_t iterations, double lambda, int64_t earlystop,
                               double* theta, double* ell, int64_t* n_done) {
  if (!h) return DSMGP_ERR_ARG;
  if (!theta || !ell || iterations <= 0 || optimiser < 0 || optimiser > 2) { h->err = "train: bad argument"; return DSMGP_ERR_ARG; }
  if (h->opts.world != 1) { h->err = "train: single-process handles only"; return DSMGP_ERR_STATE; }
  cudaSetDevice(h->device);
  const char* env = getenv("DSMGP_TRAIN_FUSED");       // "0": host loop (one dsmgp_eval per iteration), for A/B and tests
  const bool fused = !(env && env[0] == '0') && h->batches.size() == 1 && h->H <= 1024 && !h->opts.strict_pd &&
                     (double)iterations * (double)h->H < 2.5e8;
  if (fused) return train_fused(h, optimiser, eta, beta1, beta2, state_by_identity, iterations, lambda, earlystop, theta, ell, n_done);
  return train_host_loop(h, optimiser, eta, beta1, beta2, state_by_identity, iterations, lambda, earlystop, theta, ell, n_done);
}
