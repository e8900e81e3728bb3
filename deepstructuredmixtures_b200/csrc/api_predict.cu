// libdsmgp.so : prediction entry points (common.jl:101-122, 134-313; gaussianprocess.jl:110-137).
#include "handle.h"
#include "route_args.h"

using namespace dsm;
#define g_create_error (dsm::create_error())

// ------------------------------------------------------------------------------------------
// prediction
// ------------------------------------------------------------------------------------------
namespace {
struct Router {
  const HostTree& t; const double* x; int64_t T;
  std::vector<std::vector<int64_t>> pts;   // per leaf
  bool bad = false;
  Router(const HostTree& tt, const double* xx, int64_t TT, int64_t L) : t(tt), x(xx), T(TT), pts(L) {}
  void route(int64_t node, const std::vector<int64_t>& idx, bool poe) {
    const int ty = t.type[node];
    if (ty == DSMGP_NODE_LEAF) { auto& v = pts[t.leaf_of_node[node]]; v.insert(v.end(), idx.begin(), idx.end()); return; }
    if (ty == DSMGP_NODE_SPLIT && !poe) {
      std::vector<std::vector<int64_t>> sub(t.nchild(node));
      for (int64_t p : idx) { const int64_t k = getchild(t, node, x, T, p); if (k < 0) { bad = true; return; } sub[k].push_back(p); }
      for (int64_t k = 0; k < t.nchild(node); k++) if (!sub[k].empty()) route(t.child(node, k), sub[k], poe);
      return;
    }
    for (int64_t k = 0; k < t.nchild(node); k++) route(t.child(node, k), idx, poe);
  }
};

// common.jl mixing on the host.  Leaf predictions are stored per leaf in routing order; `cursor` replays it.
struct Mixer {
  const HostTree& t; const double* x; int64_t T;
  const std::vector<std::vector<double>>& mu; const std::vector<std::vector<double>>& var;
  const std::vector<double>& logw;
  std::vector<size_t> cursor;
  Mixer(const HostTree& tt, const double* xx, int64_t TT, const std::vector<std::vector<double>>& m,
        const std::vector<std::vector<double>>& v, const std::vector<double>& lw)
      : t(tt), x(xx), T(TT), mu(m), var(v), logw(lw), cursor(m.size(), 0) {}
  void reset() { std::fill(cursor.begin(), cursor.end(), 0); }

  // _minpredict common.jl:151-173
  void minpredict(int64_t node, const std::vector<int64_t>& idx, std::vector<double>& out) {
    const int ty = t.type[node];
    out.assign(idx.size(), 0.0);
    if (ty == DSMGP_NODE_LEAF) {
      const int64_t l = t.leaf_of_node[node];
      for (size_t i = 0; i < idx.size(); i++) out[i] = mu[l][cursor[l] + i];
      cursor[l] += idx.size();
    } else if (ty == DSMGP_NODE_SPLIT) {
      std::vector<std::vector<int64_t>> sub(t.nchild(node)); std::vector<std::vector<size_t>> pos(t.nchild(node));
      for (size_t i = 0; i < idx.size(); i++) { const int64_t k = getchild(t, node, x, T, idx[i]); sub[k].push_back(idx[i]); pos[k].push_back(i); }
      std::vector<double> o;
      for (int64_t k = 0; k < t.nchild(node); k++) {
        if (sub[k].empty()) continue;
        minpredict(t.child(node, k), sub[k], o);
        for (size_t i = 0; i < o.size(); i++) out[pos[k][i]] = o[i];
      }
    } else {
      std::fill(out.begin(), out.end(), std::numeric_limits<double>::infinity());
      std::vector<double> o;
      for (int64_t k = 0; k < t.nchild(node); k++) {
        minpredict(t.child(node, k), idx, o);
        for (size_t i = 0; i < o.size(); i++) out[i] = std::min(out[i], o[i]);
      }
    }
  }
  // _predict common.jl:134-143,181-196,275-292 : log(mu - mumin), log(mu^2), log(sigma^2)
  void predict(int64_t node, const std::vector<int64_t>& idx, const std::vector<double>& mumin,
               std::vector<double>& lm, std::vector<double>& lm2, std::vector<double>& ls) {
    const int ty = t.type[node];
    const size_t n = idx.size();
    lm.assign(n, 0.0); lm2.assign(n, 0.0); ls.assign(n, 0.0);
    if (ty == DSMGP_NODE_LEAF) {
      const int64_t l = t.leaf_of_node[node];
      for (size_t i = 0; i < n; i++) {
        const double m = mu[l][cursor[l] + i];
        double s2 = var[l][cursor[l] + i];
        if (s2 <= 0) s2 = 1e-8;                                  // common.jl:137
        lm[i] = std::log(m - mumin[i]); lm2[i] = std::log(m * m); ls[i] = std::log(s2);
      }
      cursor[l] += n;
    } else if (ty == DSMGP_NODE_SPLIT) {
      std::vector<std::vector<int64_t>> sub(t.nchild(node)); std::vector<std::vector<size_t>> pos(t.nchild(node));
      std::vector<std::vector<double>> mm(t.nchild(node));
      for (size_t i = 0; i < n; i++) {
        const int64_t k = getchild(t, node, x, T, idx[i]);
        sub[k].push_back(idx[i]); pos[k].push_back(i); mm[k].push_back(mumin[i]);
      }
      std::vector<double> a, b, c;
      for (int64_t k = 0; k < t.nchild(node); k++) {
        if (sub[k].empty()) continue;
        predict(t.child(node, k), sub[k], mm[k], a, b, c);
        for (size_t i = 0; i < a.size(); i++) { lm[pos[k][i]] = a[i]; lm2[pos[k][i]] = b[i]; ls[pos[k][i]] = c[i]; }
      }
    } else {
      const int64_t K = t.nchild(node);
      std::vector<std::vector<double>> A(K), B(K), C(K);
      for (int64_t k = 0; k < K; k++) predict(t.child(node, k), idx, mumin, A[k], B[k], C[k]);
      const double* lw = logw.data() + t.child_ptr[node];
      auto lse = [&](std::vector<std::vector<double>>& M, size_t i) {   // common.jl:309-313
        double m = -std::numeric_limits<double>::infinity();
        for (int64_t k = 0; k < K; k++) m = std::max(m, M[k][i] + lw[k]);
        double s = 0.0;
        for (int64_t k = 0; k < K; k++) s += std::exp((M[k][i] + lw[k]) - m);
        return std::log(s) + m;
      };
      for (size_t i = 0; i < n; i++) { lm[i] = lse(A, i); lm2[i] = lse(B, i); ls[i] = lse(C, i); }
    }
  }
  // _predictPoE common.jl:145-149,198-208 : (mu, precision)
  bool poe(int64_t node, const std::vector<int64_t>& idx, std::vector<double>& m, std::vector<double>& tau) {
    const int ty = t.type[node];
    const size_t n = idx.size();
    if (ty == DSMGP_NODE_LEAF) {
      const int64_t l = t.leaf_of_node[node];
      m.resize(n); tau.resize(n);
      for (size_t i = 0; i < n; i++) { m[i] = mu[l][cursor[l] + i]; tau[i] = 1.0 / var[l][cursor[l] + i]; }
      cursor[l] += n;
      return true;
    }
    if (ty != DSMGP_NODE_SPLIT) return false;    // MethodError in the reference
    m.assign(n, 0.0); tau.assign(n, 0.0);
    std::vector<double> m_, t_;
    for (int64_t k = 0; k < t.nchild(node); k++) {
      if (!poe(t.child(node, k), idx, m_, t_)) return false;
      for (size_t i = 0; i < n; i++) { tau[i] += t_[i]; m[i] += t_[i] * m_[i]; }
    }
    for (size_t i = 0; i < n; i++) m[i] = m[i] / tau[i];
    return true;
  }
};
}  // namespace

// Device prediction of every leaf on its routed points.  pts[l] = test rows routed to leaf l.
static int32_t predict_leaves(dsmgp_handle* h, const double* xtest, int64_t T, const std::vector<std::vector<int64_t>>& pts,
                              std::vector<std::vector<double>>& mu, std::vector<std::vector<double>>& var,
                              bool local_only = false) {
  if (!h->fitted) { h->err = "predict: call fit first"; return DSMGP_ERR_STATE; }
  if (!h->opts.keep_factors || h->batches.size() != 1) { h->err = "predict needs keep_factors=1"; return DSMGP_ERR_STATE; }
  { int32_t rr = refine_alpha(h); if (rr) return rr; }
  const int64_t L = h->L, D = h->D;
  mu.assign(L, {}); var.assign(L, {});
  std::vector<PredLeaf> pls; std::vector<int2> tasks; std::vector<int64_t> pl_leaf;
  int64_t xto = 0, vto = 0, oo = 0;
  for (int64_t l = 0; l < L; l++) {
    if (pts[l].empty()) continue;
    const int slot = h->exec_slot[l];       // an aliased expert predicts with its source's factor and alpha
    if (slot < 0) {
      if (local_only) continue;               // predicted by its owner (dsmgp_predict_local / _finish)
      h->err = "predict: leaf owned by another rank (use dsmgp_predict_local + all-reduce + dsmgp_predict_finish)";
      return DSMGP_ERR_STATE;
    }
    PredLeaf p; p.slot = slot; p.T = (int32_t)pts[l].size(); p.Tp = (p.T + BLK - 1) / BLK * BLK; p.pad_ = 0;
    p.xtoff = xto; xto += (int64_t)p.Tp * D;
    p.vtoff = vto; vto += (int64_t)(p.Tp / BLK) * h->meta[slot].nkc * TILE_D;
    p.ooff = oo; oo += p.Tp;
    for (int q = 0; q < p.Tp / BLK; q++) tasks.push_back(make_int2((int)pls.size(), q));
    pls.push_back(p); pl_leaf.push_back(l);
  }
  if (pls.empty()) return DSMGP_OK;
  std::stable_sort(tasks.begin(), tasks.end(), [&](const int2& a, const int2& b) {
    return h->meta[pls[a.x].slot].np > h->meta[pls[b.x].slot].np; });
  std::vector<double> xt(xto, 0.0);
  for (size_t i = 0; i < pls.size(); i++) {
    const auto& pv = pts[pl_leaf[i]];
    for (int64_t d = 0; d < D; d++)
      for (size_t q = 0; q < pv.size(); q++) xt[pls[i].xtoff + d * pls[i].Tp + q] = xtest[d * T + pv[q]];
  }
  DevBuf<double>&d_xt = h->p_xt, &d_VT = h->p_VT, &d_mu = h->p_mu, &d_var = h->p_var;
  DevBuf<PredLeaf>& d_pl = h->p_pl; DevBuf<int2>& d_tasks = h->p_tasks;
  auto cleanup = [&]() { if (d_VT.n * sizeof(double) > (size_t(16) << 30)) d_VT.free(); };   // keep the scratch unless it is huge
#define PTRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { h->err = std::string(#expr) + ": " + cudaGetErrorString(e_); cleanup(); \
    return e_ == cudaErrorMemoryAllocation ? DSMGP_ERR_OOM : DSMGP_ERR_CUDA; } } while (0)
  PTRY(d_xt.ensure(xto)); PTRY(d_VT.ensure(vto)); PTRY(d_mu.ensure(oo)); PTRY(d_var.ensure(oo));
  PTRY(d_pl.ensure(pls.size())); PTRY(d_tasks.ensure(tasks.size()));
  PTRY(cudaMemcpyAsync(d_xt.p, xt.data(), xto * 8, cudaMemcpyHostToDevice, h->stream));
  PTRY(cudaMemcpyAsync(d_pl.p, pls.data(), pls.size() * sizeof(PredLeaf), cudaMemcpyHostToDevice, h->stream));
  PTRY(cudaMemcpyAsync(d_tasks.p, tasks.data(), tasks.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
  PTRY(cudaMemsetAsync(h->d_counter.p, 0, 16 * sizeof(int), h->stream));
  PTRY(cudaMemsetAsync(h->d_counter.p + GERR, 0, sizeof(int), h->stream));
  PredArgs pa{h->d_meta.p, d_pl.p, d_tasks.p, (int)tasks.size(), h->d_counter.p + 2, h->d_F.p, h->d_W.p, h->d_xg.p,
              h->d_alpha.p, h->d_prm.p, h->d_leaf_mean.p, d_xt.p, d_VT.p, d_mu.p, d_var.p, (int)D, h->d_counter.p + GERR,
              0, nullptr, nullptr, nullptr, nullptr, 0};
  const int sms = num_sms(h->device);
  const char* force_wave = getenv("DSMGP_PREDICT_WAVE");       // tests: "0" / "1" force the task granularity
  const bool use_wave = force_wave ? (force_wave[0] == '1') : ((int)tasks.size() < 3 * sms);
  if (use_wave) {
    // WAVE mode: too few (leaf, Q) tasks to fill the GPU -> one task per (leaf, Q, row block), ordered by row block
    // (a block depends only on the blocks above it) with the experts shifted so that they end together
    std::vector<int4> wt, wc;
    int base = 0, max_nb = 0;
    for (auto& p : pls) max_nb = std::max(max_nb, (int)h->meta[p.slot].nb);
    struct WK { int key, np, pl, Q, I, base; };
    std::vector<WK> wk;
    for (size_t i = 0; i < pls.size(); i++) {
      const LeafMeta& m = h->meta[pls[i].slot];
      for (int q = 0; q < pls[i].Tp / BLK; q++) {
        wc.push_back(make_int4((int)i, q, base, m.nb));
        for (int I = 0; I < m.nb; I++) wk.push_back({I + max_nb - m.nb, m.np, (int)i, q, I, base});
        base += m.nb;
      }
    }
    std::stable_sort(wk.begin(), wk.end(), [](const WK& a, const WK& b) { return a.key != b.key ? a.key < b.key : a.np > b.np; });
    for (auto& k : wk) wt.push_back(make_int4(k.pl, k.Q, k.I, k.base));
    PTRY(h->p_wtasks.ensure(wt.size())); PTRY(h->p_wcols.ensure(wc.size())); PTRY(h->p_flags.ensure(base));
    PTRY(h->p_part.ensure((size_t)base * 2 * BLK));
    PTRY(cudaMemcpyAsync(h->p_wtasks.p, wt.data(), wt.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
    PTRY(cudaMemcpyAsync(h->p_wcols.p, wc.data(), wc.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
    PTRY(cudaMemsetAsync(h->p_flags.p, 0, (size_t)base * sizeof(int), h->stream));
    PTRY(cudaStreamSynchronize(h->stream));       // wt / wc are locals
    pa.wave = 1; pa.wtasks = h->p_wtasks.p; pa.ntasks = (int)wt.size(); pa.flags = h->p_flags.p; pa.part = h->p_part.p;
    pa.wcols = h->p_wcols.p; pa.nwcols = (int)wc.size();
  }
  cudaEventRecord(h->ev[0], h->stream);
  launch_predict3(pa, std::max(1, std::min(sms, pa.ntasks)), h->stream);
  if (pa.wave) launch_predict_reduce(pa, h->stream);
  cudaEventRecord(h->ev[1], h->stream);
  h->tm.launches++;
  PTRY(cudaGetLastError());
  std::vector<double> hmu(oo), hvar(oo);
  PTRY(cudaMemcpyAsync(hmu.data(), d_mu.p, oo * 8, cudaMemcpyDeviceToHost, h->stream));
  PTRY(cudaMemcpyAsync(hvar.data(), d_var.p, oo * 8, cudaMemcpyDeviceToHost, h->stream));
  int gerr = 0;
  PTRY(cudaMemcpyAsync(&gerr, h->d_counter.p + GERR, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  PTRY(cudaStreamSynchronize(h->stream));
  if (gerr != 0) { h->err = "predict: device scheduler timeout (code " + std::to_string(gerr) + ")"; cleanup(); return DSMGP_ERR_STATE; }
#undef PTRY
  h->tm.predict_ms = ev_ms(h->ev[0], h->ev[1]);
  h->tm.predict_flops = 0.0; h->tm.predict_bytes = 0.0;
  for (size_t i = 0; i < pls.size(); i++) {      // TRSM n^2 T_l + 2 n T_l flop; L read once per block of 128 points
    const double n = h->meta[pls[i].slot].n, Tl = pls[i].T;
    h->tm.predict_flops += n * n * Tl + 2.0 * n * Tl;
    h->tm.predict_bytes += 8.0 * (n * (n + 1) / 2.0) * (pls[i].Tp / BLK) + 8.0 * (n + Tl) * (double)D + 16.0 * Tl;
  }
  for (size_t i = 0; i < pls.size(); i++) {
    const int64_t l = pl_leaf[i];
    mu[l].assign(hmu.begin() + pls[i].ooff, hmu.begin() + pls[i].ooff + pls[i].T);
    var[l].assign(hvar.begin() + pls[i].ooff, hvar.begin() + pls[i].ooff + pls[i].T);
  }
  cleanup();
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_leaf_predict(dsmgp_handle* h, int64_t leaf, const double* xtest, int64_t T, double* mu, double* var) {
  if (!h || leaf < 0 || leaf >= h->L || !xtest || T <= 0 || !mu || !var) return DSMGP_ERR_ARG;
  cudaSetDevice(h->device);
  std::vector<std::vector<int64_t>> pts(h->L);
  pts[leaf].resize(T);
  std::iota(pts[leaf].begin(), pts[leaf].end(), 0);
  std::vector<std::vector<double>> m, v;
  int32_t rc = predict_leaves(h, xtest, T, pts, m, v);
  if (rc) return rc;
  std::copy(m[leaf].begin(), m[leaf].end(), mu);
  std::copy(v[leaf].begin(), v[leaf].end(), var);
  return DSMGP_OK;
}

// argument checks + routing shared by the predict entry points
static int32_t predict_route(dsmgp_handle* h, const double* xtest, int64_t T, int32_t mode, std::vector<std::vector<int64_t>>& pts) {
  if (!xtest || T <= 0 || mode < 0 || mode > 3) { h->err = "predict: bad argument"; return DSMGP_ERR_ARG; }
  for (int64_t i = 0; i < T * h->D; i++) if (!std::isfinite(xtest[i])) { h->err = "predict: non-finite input"; return DSMGP_ERR_ARG; }
  cudaSetDevice(h->device);
  const HostTree& t = h->tree;
  const bool poe = mode != DSMGP_PREDICT_DSMGP;
  if (poe && t.type[t.root] != DSMGP_NODE_SPLIT) { h->err = "predict: PoE/gPoE/rBCM need a split root (buildPoE/buildBCM model)"; return DSMGP_ERR_ARG; }
  if (!poe && !h->have_weights) {
    // the reference predicts with whatever logweights the sum nodes hold (uniform -log K after build)
    h->sum_logw.assign(t.child_ptr[t.n_nodes], 0.0);
    for (int64_t i = 0; i < t.n_nodes; i++)
      if (t.type[i] >= DSMGP_NODE_SUM) for (int64_t c = t.child_ptr[i]; c < t.child_ptr[i + 1]; c++) h->sum_logw[c] = -std::log((double)t.nchild(i));
  }
  std::vector<int64_t> all(T);
  std::iota(all.begin(), all.end(), 0);
  Router r(t, xtest, T, h->L);
  r.route(t.root, all, poe);
  if (r.bad) { h->err = "predict: a test point lies outside every split interval"; return DSMGP_ERR_ARG; }
  pts.swap(r.pts);
  return DSMGP_OK;
}

static int32_t predict_mix(dsmgp_handle* h, const double* xtest, int64_t T, int32_t mode,
                           const std::vector<std::vector<double>>& lmu, const std::vector<std::vector<double>>& lvar,
                           double* mu, double* var);

// ---- large batches: route, predict and mix on the device ------------------------------------------------------------------
// The host path above walks the tree with std::vector recursion, packs the routed points per expert, uploads them and mixes
// on the host: for 40,000 points that costs 50 ms next to a 130 ms kernel.  Here x_test is uploaded ONCE, one thread per point
// routes it (route.cuh), predict3 gathers its test tiles through the index lists, and one thread per point mixes
// (common.jl:134-313).  Only the per-expert counts (L ints) visit the host in between, to size the task list.
static int tree_metrics(const HostTree& t, int64_t node, bool poe, int depth, int* maxdepth, int* maxk, int frames, int* maxframes) {
  *maxdepth = std::max(*maxdepth, depth);
  const int ty = t.type[node];
  if (ty == DSMGP_NODE_LEAF) { *maxframes = std::max(*maxframes, frames); return 1; }
  *maxk = std::max(*maxk, (int)t.nchild(node));
  const bool fan = (ty != DSMGP_NODE_SPLIT) || poe;            // every child is visited
  const int fr = frames + ((ty != DSMGP_NODE_SPLIT) != poe ? 1 : 0);   // DSMGP: frames are sum nodes; PoE: split nodes
  int reach = 0;
  for (int64_t k = 0; k < t.nchild(node); k++) {
    const int r = tree_metrics(t, t.child(node, k), poe, depth + 1, maxdepth, maxk, fr, maxframes);
    reach = fan ? reach + r : std::max(reach, r);
  }
  return reach;
}

int32_t dsm::ensure_dev_tree(dsmgp_handle* h) {
  if (h->tree_on_device) return DSMGP_OK;
  const HostTree& t = h->tree;
  const int64_t nn = t.n_nodes, nc = t.child_ptr[nn], nsv = t.split_ptr[nn];
  std::vector<int> buf;
  auto push = [&](auto& v, int64_t n) { for (int64_t i = 0; i < n; i++) buf.push_back((int)v[i]); };
  push(t.type, nn); push(t.child_ptr, nn + 1); push(t.child_idx, nc); push(t.leaf_of_node, nn); push(t.split_dim, nn); push(t.split_ptr, nn + 1);
  CUDA_TRY(h, h->t_int.alloc(buf.size()));
  CUDA_TRY(h, cudaMemcpy(h->t_int.p, buf.data(), buf.size() * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_TRY(h, h->t_split_val.alloc(std::max<int64_t>(nsv, 1)));
  if (nsv) CUDA_TRY(h, cudaMemcpy(h->t_split_val.p, t.split_val.data(), nsv * sizeof(double), cudaMemcpyHostToDevice));
  int md = 0, mk = 0, mf = 0, md2 = 0, mk2 = 0, mf2 = 0;
  h->reach_dsmgp = tree_metrics(t, t.root, false, 1, &md, &mk, 0, &mf);
  h->reach_poe = tree_metrics(t, t.root, true, 1, &md2, &mk2, 0, &mf2);
  h->tree_depth = md; h->tree_maxk = mk; h->tree_frames = std::max(mf, mf2);
  h->tree_on_device = true;
  return DSMGP_OK;
}

DevTree dsm::dev_tree(const dsmgp_handle* h) {
  const HostTree& t = h->tree;
  const int64_t nn = t.n_nodes, nc = t.child_ptr[nn];
  DevTree d;
  d.n_nodes = (int)nn; d.root = (int)t.root;
  const int* p = h->t_int.p;
  d.type = p; p += nn; d.child_ptr = p; p += nn + 1; d.child_idx = p; p += nc; d.leaf_of_node = p; p += nn; d.split_dim = p; p += nn; d.split_ptr = p;
  d.split_val = h->t_split_val.p;
  return d;
}

// Can (and should) this prediction run on the device path?
static bool predict_on_device(dsmgp_handle* h, int64_t T, int32_t mode) {
  if (h->opts.world != 1 || !h->fitted || !h->opts.keep_factors || h->batches.size() != 1) return false;
  const char* force = getenv("DSMGP_PREDICT_DEVICE");       // tests: "0" / "1" force the path
  if (force && force[0] == '0') return false;
  if (ensure_dev_tree(h) != DSMGP_OK) return false;
  const bool poe = mode != DSMGP_PREDICT_DSMGP;
  const int R = poe ? h->reach_poe : h->reach_dsmgp;
  if (h->tree_maxk > MIX_KMAX || h->tree_frames > MIX_FRAMES || h->tree_depth * std::max(h->tree_maxk, 1) > ROUTE_STACK) return false;
  if ((double)T * R > 1.5e9) return false;
  if (poe) for (int64_t i = 0; i < h->tree.n_nodes; i++) if (h->tree.type[i] >= DSMGP_NODE_SUM) return false;   // MethodError path: host reports it
  if (force && force[0] == '1') return true;
  return T * R >= 3ll * num_sms(h->device) * BLK;           // enough (expert, 128-point block) tasks to fill the GPU
}

static int32_t predict_device(dsmgp_handle* h, const double* xtest, int64_t T, int32_t mode, double* mu, double* var) {
  cudaStream_t st = h->stream;
  const HostTree& t = h->tree;
  const int64_t L = h->L, D = h->D;
  const bool poe = mode != DSMGP_PREDICT_DSMGP;
  if (poe && t.type[t.root] != DSMGP_NODE_SPLIT) { h->err = "predict: PoE/gPoE/rBCM need a split root (buildPoE/buildBCM model)"; return DSMGP_ERR_ARG; }
  { int32_t rr = refine_alpha(h); if (rr) return rr; }
  if (!poe && !h->have_weights) {       // the reference predicts with whatever logweights the sum nodes hold (uniform after build)
    h->sum_logw.assign(t.child_ptr[t.n_nodes], 0.0);
    reset_weights(t, h->sum_logw.data());
  }
  const int R = poe ? h->reach_poe : h->reach_dsmgp;
  CUDA_TRY(h, h->p_xtest.ensure((size_t)T * D));
  CUDA_TRY(h, cudaMemcpyAsync(h->p_xtest.p, xtest, (size_t)T * D * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, h->p_cnt.ensure(3 * L + 1));
  CUDA_TRY(h, cudaMemsetAsync(h->p_cnt.p, 0, (3 * L + 1) * sizeof(int), st));
  CUDA_TRY(h, h->p_reach.ensure((size_t)T * R));
  CUDA_TRY(h, cudaMemsetAsync(h->p_reach.p, 0xFF, (size_t)T * R * sizeof(int), st));
  RouteArgs ra{dev_tree(h), h->p_xtest.p, T, (int)D, poe ? 1 : 0, R, h->p_cnt.p, h->p_cnt.p + 2 * L, h->p_cnt.p + L, nullptr,
               h->p_reach.p, h->p_cnt.p + 3 * L};
  launch_route(ra, false, st);
  std::vector<int> cnt(3 * L + 1);
  CUDA_TRY(h, cudaMemcpyAsync(cnt.data(), h->p_cnt.p, (3 * L + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(h, cudaStreamSynchronize(st));
  if (cnt[3 * L] == 1) { h->err = "predict: non-finite input"; return DSMGP_ERR_ARG; }
  if (cnt[3 * L] == 2) { h->err = "predict: a test point lies outside every split interval"; return DSMGP_ERR_ARG; }
  // experts with points, largest first (slots are sorted by size); outputs of an expert are padded to whole 128-point blocks
  std::vector<int64_t> order;
  for (int64_t l = 0; l < L; l++) if (cnt[l] > 0) {
    if (h->exec_slot[l] < 0) { h->err = "predict: leaf owned by another rank"; return DSMGP_ERR_STATE; }
    order.push_back(l);
  }
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return h->meta[h->exec_slot[a]].np > h->meta[h->exec_slot[b]].np; });
  std::vector<PredLeaf> pls; std::vector<int2> tasks; std::vector<int> ooff(L, 0);
  int64_t oo = 0; int max_nkc = 0;
  for (int64_t l : order) {
    PredLeaf pl; pl.slot = h->exec_slot[l]; pl.T = cnt[l]; pl.Tp = (pl.T + BLK - 1) / BLK * BLK; pl.pad_ = 0;
    pl.xtoff = 0; pl.vtoff = 0; pl.ooff = oo;
    ooff[l] = (int)oo; oo += pl.Tp;
    max_nkc = std::max(max_nkc, (int)h->meta[pl.slot].nkc);
    for (int q = 0; q < pl.Tp / BLK; q++) tasks.push_back(make_int2((int)pls.size(), q));
    pls.push_back(pl);
  }
  if (oo >= (int64_t(1) << 31)) { h->err = "predict: too many (expert, point) pairs for the device path"; return DSMGP_ERR_ARG; }
  const int sms = num_sms(h->device);
  const int nctas = std::max(1, std::min(sms, (int)tasks.size()));
  const int64_t vt_stride = (int64_t)max_nkc * TILE_D;
  CUDA_TRY(h, h->p_pl.ensure(pls.size())); CUDA_TRY(h, h->p_tasks.ensure(tasks.size()));
  CUDA_TRY(h, h->p_pidx.ensure(oo)); CUDA_TRY(h, h->p_mu.ensure(oo)); CUDA_TRY(h, h->p_var.ensure(oo));
  CUDA_TRY(h, h->p_VT.ensure((size_t)nctas * vt_stride));
  CUDA_TRY(h, h->p_outmu.ensure(T)); CUDA_TRY(h, h->p_outvar.ensure(T));
  CUDA_TRY(h, h->p_logw.ensure(std::max<size_t>(h->sum_logw.size(), 1)));
  CUDA_TRY(h, cudaMemcpyAsync(h->p_pl.p, pls.data(), pls.size() * sizeof(PredLeaf), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaMemcpyAsync(h->p_tasks.p, tasks.data(), tasks.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaMemcpyAsync(h->p_cnt.p + 2 * L, ooff.data(), L * sizeof(int), cudaMemcpyHostToDevice, st));
  if (!h->sum_logw.empty()) CUDA_TRY(h, cudaMemcpyAsync(h->p_logw.p, h->sum_logw.data(), h->sum_logw.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaMemsetAsync(h->p_pidx.p, 0xFF, oo * sizeof(int), st));
  ra.pidx = h->p_pidx.p;
  launch_route(ra, true, st);
  CUDA_TRY(h, cudaMemsetAsync(h->d_counter.p, 0, 16 * sizeof(int), st));
  CUDA_TRY(h, cudaMemsetAsync(h->d_counter.p + GERR, 0, sizeof(int), st));
  PredArgs pa{h->d_meta.p, h->p_pl.p, h->p_tasks.p, (int)tasks.size(), h->d_counter.p + 2, h->d_F.p, h->d_W.p, h->d_xg.p,
              h->d_alpha.p, h->d_prm.p, h->d_leaf_mean.p, nullptr, h->p_VT.p, h->p_mu.p, h->p_var.p, (int)D, h->d_counter.p + GERR,
              0, nullptr, nullptr, nullptr, nullptr, 0, h->p_pidx.p, h->p_xtest.p, T, 1, vt_stride, nullptr};
  long long* d_trace = nullptr;
  const char* trace_file = getenv("DSMGP_PTRACE_FILE");       // development: per-task phase cycle counts (tools/trace_predict.py)
  if (trace_file) { cudaMalloc(&d_trace, tasks.size() * 64); cudaMemsetAsync(d_trace, 0, tasks.size() * 64, st); pa.trace = d_trace; }
  cudaEventRecord(h->ev[0], st);
  launch_predict3(pa, nctas, st);
  cudaEventRecord(h->ev[1], st);
  if (trace_file) {
    std::vector<long long> tr(tasks.size() * 8);
    cudaMemcpyAsync(tr.data(), d_trace, tr.size() * 8, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    if (FILE* f = fopen(trace_file, "wb")) { fwrite(tr.data(), 8, tr.size(), f); fclose(f); }
    cudaFree(d_trace);
  }
  h->tm.launches = 4;
  // rBCM prior: the left-most expert's kernel (common.jl:226-227)
  int64_t nd = t.root;
  while (t.type[nd] != DSMGP_NODE_LEAF) nd = t.child(nd, 0);
  const int s0 = h->exec_slot[t.leaf_of_node[nd]];
  MixArgs ma{dev_tree(h), h->p_xtest.p, T, (int)D, mode, R, h->p_reach.p, h->p_mu.p, h->p_var.p, h->p_logw.p,
             s0 >= 0 ? h->meta[s0].ktype : 0, s0 >= 0 ? h->d_prm.p + h->meta[s0].poff : h->d_prm.p, h->p_outmu.p, h->p_outvar.p};
  launch_mix(ma, st);
  CUDA_TRY(h, cudaGetLastError());
  CUDA_TRY(h, cudaMemcpyAsync(mu, h->p_outmu.p, T * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(h, cudaMemcpyAsync(var, h->p_outvar.p, T * sizeof(double), cudaMemcpyDeviceToHost, st));
  int gerr = 0;
  CUDA_TRY(h, cudaMemcpyAsync(&gerr, h->d_counter.p + GERR, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(h, cudaStreamSynchronize(st));
  if (gerr != 0) { h->err = "predict: device scheduler timeout (code " + std::to_string(gerr) + ")"; return DSMGP_ERR_STATE; }
  h->tm.predict_ms = ev_ms(h->ev[0], h->ev[1]);
  h->tm.predict_flops = 0.0; h->tm.predict_bytes = 0.0;
  for (auto& pl : pls) {
    const double n = h->meta[pl.slot].n, Tl = pl.T;
    h->tm.predict_flops += n * n * Tl + 2.0 * n * Tl;
    h->tm.predict_bytes += 8.0 * (n * (n + 1) / 2.0) * (pl.Tp / BLK) + 8.0 * (n + Tl) * (double)D + 16.0 * Tl;
  }
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_predict(dsmgp_handle* h, const double* xtest, int64_t T, int32_t mode, double* mu, double* var) {
  if (!h) return DSMGP_ERR_ARG;
  if (!mu || !var) { h->err = "predict: bad argument"; return DSMGP_ERR_ARG; }
  if (xtest && T > 0 && mode >= 0 && mode <= 3) {
    cudaSetDevice(h->device);
    if (predict_on_device(h, T, mode)) return predict_device(h, xtest, T, mode, mu, var);
  }
  std::vector<std::vector<int64_t>> pts;
  int32_t rc = predict_route(h, xtest, T, mode, pts);
  if (rc) return rc;
  std::vector<std::vector<double>> lmu, lvar;
  if (h->opts.world > 1) {
    // experts sharded over ranks: predict the local ones, assemble the (expert, point) table with ONE SUM all-reduce on the
    // library's communicator, mix on every rank
    if (!h->comm) { h->err = "predict: world > 1 needs dsmgp_comm_init (or dsmgp_predict_local + all-reduce + dsmgp_predict_finish)"; return DSMGP_ERR_STATE; }
    if ((rc = predict_leaves(h, xtest, T, pts, lmu, lvar, true))) return rc;
    int64_t tot = 0;
    for (auto& v : pts) tot += (int64_t)v.size();
    std::vector<double> buf(2 * tot, 0.0);
    int64_t off = 0;
    for (int64_t l = 0; l < h->L; l++) {
      if (!lmu.empty() && !lmu[l].empty()) {
        std::copy(lmu[l].begin(), lmu[l].end(), buf.begin() + off);
        std::copy(lvar[l].begin(), lvar[l].end(), buf.begin() + tot + off);
      }
      off += (int64_t)pts[l].size();
    }
    CUDA_TRY(h, h->d_comm_buf.ensure(2 * tot));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_comm_buf.p, buf.data(), 2 * tot * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if ((rc = comm_allreduce_sum(h, h->d_comm_buf.p, 2 * tot))) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(buf.data(), h->d_comm_buf.p, 2 * tot * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    lmu.assign(h->L, {}); lvar.assign(h->L, {});
    off = 0;
    for (int64_t l = 0; l < h->L; l++) {
      lmu[l].assign(buf.begin() + off, buf.begin() + off + pts[l].size());
      lvar[l].assign(buf.begin() + tot + off, buf.begin() + tot + off + pts[l].size());
      off += (int64_t)pts[l].size();
    }
    return predict_mix(h, xtest, T, mode, lmu, lvar, mu, var);
  }
  if ((rc = predict_leaves(h, xtest, T, pts, lmu, lvar))) return rc;
  return predict_mix(h, xtest, T, mode, lmu, lvar, mu, var);
}

// Leaf-sharded prediction (one process per GPU): every rank predicts its own experts on the points routed to them and
// writes them into a buffer in (leaf, routing order) layout -- entries of other ranks' experts stay 0, so a SUM all-reduce
// assembles the buffer -- then every rank mixes (common.jl:134-307) redundantly, like the tree passes of an evaluation.
extern "C" int32_t dsmgp_predict_local(dsmgp_handle* h, const double* xtest, int64_t T, int32_t mode, double* buf, int64_t* total) {
  if (!h) return DSMGP_ERR_ARG;
  if (!total) { h->err = "predict_local: bad argument"; return DSMGP_ERR_ARG; }
  std::vector<std::vector<int64_t>> pts;
  int32_t rc = predict_route(h, xtest, T, mode, pts);
  if (rc) return rc;
  int64_t tot = 0;
  for (auto& v : pts) tot += (int64_t)v.size();
  *total = tot;
  if (!buf) return DSMGP_OK;                   // size query
  std::vector<std::vector<double>> lmu, lvar;
  if ((rc = predict_leaves(h, xtest, T, pts, lmu, lvar, true))) return rc;
  std::fill(buf, buf + 2 * tot, 0.0);
  int64_t off = 0;
  for (int64_t l = 0; l < h->L; l++) {
    if (!lmu.empty() && !lmu[l].empty()) {
      std::copy(lmu[l].begin(), lmu[l].end(), buf + off);
      std::copy(lvar[l].begin(), lvar[l].end(), buf + tot + off);
    }
    off += (int64_t)pts[l].size();
  }
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_predict_finish(dsmgp_handle* h, const double* xtest, int64_t T, int32_t mode, const double* buf,
                                        double* mu, double* var) {
  if (!h) return DSMGP_ERR_ARG;
  if (!buf || !mu || !var) { h->err = "predict_finish: bad argument"; return DSMGP_ERR_ARG; }
  std::vector<std::vector<int64_t>> pts;
  int32_t rc = predict_route(h, xtest, T, mode, pts);
  if (rc) return rc;
  int64_t tot = 0;
  for (auto& v : pts) tot += (int64_t)v.size();
  std::vector<std::vector<double>> lmu(h->L), lvar(h->L);
  int64_t off = 0;
  for (int64_t l = 0; l < h->L; l++) {
    lmu[l].assign(buf + off, buf + off + pts[l].size());
    lvar[l].assign(buf + tot + off, buf + tot + off + pts[l].size());
    off += (int64_t)pts[l].size();
  }
  return predict_mix(h, xtest, T, mode, lmu, lvar, mu, var);
}

static int32_t predict_mix(dsmgp_handle* h, const double* xtest, int64_t T, int32_t mode,
                           const std::vector<std::vector<double>>& lmu, const std::vector<std::vector<double>>& lvar,
                           double* mu, double* var) {
  const HostTree& t = h->tree;
  std::vector<int64_t> all(T);
  std::iota(all.begin(), all.end(), 0);
  Mixer mx(t, xtest, T, lmu, lvar, h->sum_logw);
  if (mode == DSMGP_PREDICT_DSMGP) {
    // predict(node) common.jl:175-179 (leaf), :243-254 (split root), :294-302 (sum root)
    struct Rec {
      Mixer& mx; const HostTree& t; double* mu; double* var; const double* x; int64_t T;
      void run(int64_t node, const std::vector<int64_t>& idx) {
        if (t.type[node] == DSMGP_NODE_SPLIT) {
          std::vector<std::vector<int64_t>> sub(t.nchild(node));
          for (int64_t p : idx) sub[getchild(t, node, x, T, p)].push_back(p);
          for (int64_t k = 0; k < t.nchild(node); k++) if (!sub[k].empty()) run(t.child(node, k), sub[k]);
          return;
        }
        // leaf or sum: two traversals of the subtree -> replay cursors must restart for this subtree.
        std::vector<size_t> save = mx.cursor;
        std::vector<double> mumin, lm, lm2, ls;
        mx.minpredict(node, idx, mumin);
        mx.cursor = save;
        for (auto& v : mumin) v -= 1.0;
        mx.predict(node, idx, mumin, lm, lm2, ls);
        for (size_t i = 0; i < idx.size(); i++) {
          const double m = std::exp(lm[i]) + mumin[i];
          mu[idx[i]] = m;
          var[idx[i]] = (t.type[node] == DSMGP_NODE_LEAF) ? std::exp(ls[i]) : std::exp(ls[i]) + (std::exp(lm2[i]) - m * m);
        }
      }
    } rec{mx, t, mu, var, xtest, T};
    rec.run(t.root, all);
    return DSMGP_OK;
  }
  const int64_t K = t.nchild(t.root);
  std::vector<double> m_, t_;
  if (mode == DSMGP_PREDICT_POE) {
    if (!mx.poe(t.root, all, m_, t_)) { h->err = "predictPoE: sum node below a split (MethodError in the reference)"; return DSMGP_ERR_ARG; }
    for (int64_t i = 0; i < T; i++) { mu[i] = m_[i]; var[i] = 1.0 / t_[i]; }
  } else if (mode == DSMGP_PREDICT_GPOE) {       // common.jl:211-222
    const double beta = 1.0 / (double)K;
    std::vector<double> M(T, 0.0), Tt(T, 0.0);
    for (int64_t k = 0; k < K; k++) {
      if (!mx.poe(t.child(t.root, k), all, m_, t_)) { h->err = "predictgPoE: sum node below a split"; return DSMGP_ERR_ARG; }
      for (int64_t i = 0; i < T; i++) { Tt[i] += beta * t_[i]; M[i] += beta * t_[i] * m_[i]; }
    }
    for (int64_t i = 0; i < T; i++) { mu[i] = M[i] / Tt[i]; var[i] = 1.0 / Tt[i]; }
  } else {                                       // rBCM common.jl:224-241
    int64_t nd = t.root;
    while (t.type[nd] != DSMGP_NODE_LEAF) nd = t.child(nd, 0);
    const int64_t l0 = t.leaf_of_node[nd];
    const int k0 = h->leaf_kid[l0];
    std::vector<double> prm(h->pstride);
    derive_params(h, k0, &h->theta_leaf[(size_t)l0 * h->Hmax], prm.data());
    const int type = h->kernels[k0].type;
    std::vector<double> s(T), C(T), M(T, 0.0);
    for (int64_t i = 0; i < T; i++) {
      double ktt;
      if (type == DSMGP_ISO_SE) ktt = prm[PRM_V];
      else if (type == DSMGP_ARD_SE) ktt = prm[PRM_V] * (double)h->D;
      else {
        ktt = 0.0;
        for (int64_t d = 0; d < h->D; d++) { const double xv = xtest[d * T + i]; ktt += (type == DSMGP_ISO_LINEAR ? prm[PRM_COEF] : prm[PRM_COEF + d]) * xv * xv; }
      }
      s[i] = ktt + prm[PRM_ETA];
      C[i] = 1.0 / s[i];
    }
    for (int64_t k = 0; k < K; k++) {
      if (!mx.poe(t.child(t.root, k), all, m_, t_)) { h->err = "predictrBCM: sum node below a split"; return DSMGP_ERR_ARG; }
      for (int64_t i = 0; i < T; i++) {
        const double s_ = 1.0 / t_[i];
        const double beta = 0.5 * (std::log(s[i]) - std::log(s_));
        C[i] = C[i] + (beta * t_[i]) - (beta / s[i]);
        M[i] = M[i] + m_[i] * (beta * t_[i]);
      }
    }
    for (int64_t i = 0; i < T; i++) { mu[i] = M[i] / C[i]; var[i] = 1.0 / C[i]; }
  }
  return DSMGP_OK;
}
