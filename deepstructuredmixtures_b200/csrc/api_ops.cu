// libdsmgp.so : stand-alone operators (kernelmatrix, getOverlap, potrf / chol_continue / row deletion on one matrix).
#include "handle.h"

using namespace dsm;
#define g_create_error (dsm::create_error())

// ------------------------------------------------------------------------------------------
// stand-alone operators
// ------------------------------------------------------------------------------------------
int32_t dsm::standalone_device_check(std::string& err) {
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) { err = std::string("no CUDA device: libdsmgp has no CPU fallback (") + cudaGetErrorString(ce) + ")"; return DSMGP_ERR_CUDA; }
  if ((ce = engine_attrs()) != cudaSuccess) { err = cudaGetErrorString(ce); return DSMGP_ERR_CUDA; }
  return DSMGP_OK;
}
#define SA_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { g_create_error = std::string(#expr) + ": " + cudaGetErrorString(e_); rc = DSMGP_ERR_CUDA; goto done; } } while (0)

extern "C" int32_t dsmgp_kernelmatrix(int32_t kernel_type, const double* theta, int64_t D, const double* x1, int64_t n1,
                                      const double* x2, int64_t n2, double* K) {
  if (kernel_type < 0 || kernel_type > 3 || !theta || !x1 || !x2 || !K || D <= 0 || n1 <= 0 || n2 <= 0) return DSMGP_ERR_ARG;
  int32_t rc = standalone_device_check(g_create_error);
  if (rc) return rc;
  const bool iso = (kernel_type == DSMGP_ISO_SE || kernel_type == DSMGP_ISO_LINEAR);
  const bool se = (kernel_type == DSMGP_ISO_SE || kernel_type == DSMGP_ARD_SE);
  const int nl = iso ? 1 : (int)D;
  std::vector<double> prm(PRM_COEF + D, 0.0);
  prm[PRM_V] = se ? std::exp(2.0 * theta[nl]) : 1.0;
  prm[PRM_S] = se ? std::exp(theta[nl]) : 1.0;
  for (int d = 0; d < nl; d++) { const double l = std::exp(theta[d]); prm[PRM_COEF + d] = se ? -0.5 / (l * l) : 1.0 / (l * l); }
  double *d1 = nullptr, *d2 = nullptr, *dk = nullptr, *dp = nullptr;
  SA_TRY(cudaMalloc(&d1, n1 * D * 8)); SA_TRY(cudaMalloc(&d2, n2 * D * 8)); SA_TRY(cudaMalloc(&dk, n1 * n2 * 8));
  SA_TRY(cudaMalloc(&dp, prm.size() * 8));
  SA_TRY(cudaMemcpy(d1, x1, n1 * D * 8, cudaMemcpyHostToDevice));
  SA_TRY(cudaMemcpy(d2, x2, n2 * D * 8, cudaMemcpyHostToDevice));
  SA_TRY(cudaMemcpy(dp, prm.data(), prm.size() * 8, cudaMemcpyHostToDevice));
  {
    GramRectArgs ga{kernel_type, (int)D, dp, d1, n1, (int)n1, d2, n2, (int)n2, dk, n1};
    launch_gram_rect(ga, 0);
    SA_TRY(cudaGetLastError());
    SA_TRY(cudaMemcpy(K, dk, n1 * n2 * 8, cudaMemcpyDeviceToHost));
  }
done:
  cudaFree(d1); cudaFree(d2); cudaFree(dk); cudaFree(dp);
  return rc;
}

// potrf / chol_continue on one host matrix.  k = number of leading rows/cols that already hold a valid factor.
// Runs the same persistent tile scheduler as the batched path (potrf2_kernel) on a one-expert batch.
// getOverlap(spn, D, gpmap) fit.jl:12-39 on the device.
// dense == true: D is the L x L column-major matrix.  dense == false: CSR by row (row_ptr[L+1] always; col / val when non-null).
static int32_t overlap_core(int64_t N, int64_t L, const int64_t* leaf_ptr, const int64_t* leaf_obs, const int32_t* leaf_kernel_id,
                            const dsmgp_tree* tree, bool dense, double* D, int64_t* row_ptr, int32_t* col, double* val) {
  if (N <= 0 || L <= 0 || !leaf_ptr || !leaf_obs || !leaf_kernel_id || !tree || (dense && !D) || (!dense && !row_ptr)) { g_create_error = "overlap: bad argument"; return DSMGP_ERR_ARG; }
  HostTree t; std::string err;
  if (!t.load(tree, L, err)) { g_create_error = err; return DSMGP_ERR_ARG; }
  const int64_t total = leaf_ptr[L];
  for (int64_t i = 0; i < total; i++) if (leaf_obs[i] < 1 || leaf_obs[i] > N) { g_create_error = "overlap: leaf_obs must be 1-based rows in 1..N"; return DSMGP_ERR_ARG; }
  { int32_t rc = standalone_device_check(g_create_error); if (rc) return rc; }
  // ancestor chains (root first) from the child lists
  std::vector<int64_t> parent(t.n_nodes, -1);
  for (int64_t i = 0; i < t.n_nodes; i++) for (int64_t k = 0; k < t.nchild(i); k++) parent[t.child(i, k)] = i;
  std::vector<std::vector<int>> chain(L);
  int AD = 1;
  for (int64_t i = 0; i < t.n_nodes; i++) {
    if (t.type[i] != DSMGP_NODE_LEAF) continue;
    std::vector<int> c;
    for (int64_t u = parent[i]; u >= 0; u = parent[u]) c.push_back((int)u);
    std::reverse(c.begin(), c.end());
    AD = std::max<int>(AD, (int)c.size());
    chain[t.leaf_of_node[i]] = c;
  }
  std::vector<int> anc((size_t)L * AD, -1);
  for (int64_t l = 0; l < L; l++) std::copy(chain[l].begin(), chain[l].end(), anc.begin() + (size_t)l * AD);
  std::vector<int> ntype(t.type.begin(), t.type.end()), kid(leaf_kernel_id, leaf_kernel_id + L);
  int64_t* d_obs = nullptr; int64_t* d_lp = nullptr; int64_t* d_poff = nullptr;
  int *d_cnt = nullptr, *d_plist = nullptr, *d_inter = nullptr, *d_kid = nullptr, *d_anc = nullptr, *d_nt = nullptr;
  double* d_D = nullptr;
  auto cleanup = [&]() { cudaFree(d_obs); cudaFree(d_lp); cudaFree(d_poff); cudaFree(d_cnt); cudaFree(d_plist); cudaFree(d_inter);
                         cudaFree(d_kid); cudaFree(d_anc); cudaFree(d_nt); cudaFree(d_D); };
#define OTRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { g_create_error = std::string(#expr) + ": " + cudaGetErrorString(e_); cleanup(); \
    return e_ == cudaErrorMemoryAllocation ? DSMGP_ERR_OOM : DSMGP_ERR_CUDA; } } while (0)
  OTRY(cudaMalloc(&d_obs, total * 8)); OTRY(cudaMemcpy(d_obs, leaf_obs, total * 8, cudaMemcpyHostToDevice));
  OTRY(cudaMalloc(&d_lp, (L + 1) * 8)); OTRY(cudaMemcpy(d_lp, leaf_ptr, (L + 1) * 8, cudaMemcpyHostToDevice));
  OTRY(cudaMalloc(&d_cnt, N * 4)); OTRY(cudaMemset(d_cnt, 0, N * 4));
  launch_ov_count(d_obs, total, d_cnt, nullptr);
  std::vector<int> cnt(N);
  OTRY(cudaMemcpy(cnt.data(), d_cnt, N * 4, cudaMemcpyDeviceToHost));
  std::vector<int64_t> poff(N + 1, 0);
  for (int64_t p = 0; p < N; p++) poff[p + 1] = poff[p] + cnt[p];
  OTRY(cudaMalloc(&d_poff, (N + 1) * 8)); OTRY(cudaMemcpy(d_poff, poff.data(), (N + 1) * 8, cudaMemcpyHostToDevice));
  OTRY(cudaMalloc(&d_plist, std::max<int64_t>(total, 1) * 4));
  OTRY(cudaMemset(d_cnt, 0, N * 4));
  launch_ov_fill(d_obs, d_lp, (int)L, d_poff, d_cnt, d_plist, nullptr);
  OTRY(cudaMalloc(&d_inter, (size_t)L * L * 4)); OTRY(cudaMemset(d_inter, 0, (size_t)L * L * 4));
  launch_ov_pairs(d_poff, d_plist, N, L, d_inter, nullptr);
  OTRY(cudaMalloc(&d_kid, L * 4)); OTRY(cudaMemcpy(d_kid, kid.data(), L * 4, cudaMemcpyHostToDevice));
  OTRY(cudaMalloc(&d_anc, anc.size() * 4)); OTRY(cudaMemcpy(d_anc, anc.data(), anc.size() * 4, cudaMemcpyHostToDevice));
  OTRY(cudaMalloc(&d_nt, ntype.size() * 4)); OTRY(cudaMemcpy(d_nt, ntype.data(), ntype.size() * 4, cudaMemcpyHostToDevice));
  if (dense) {
    OTRY(cudaMalloc(&d_D, (size_t)L * L * 8));
    launch_ov_finish(d_inter, d_lp, d_kid, d_anc, AD, d_nt, L, d_D, nullptr);
    OTRY(cudaGetLastError());
    OTRY(cudaMemcpy(D, d_D, (size_t)L * L * 8, cudaMemcpyDeviceToHost));
  } else {
    // sparse: only the non-zero entries leave the device (experts under different children of a common sum node that share
    // observations, or differ in kernel id) -- the dense L x L double matrix (3.4 GB at L = 20,736) is never formed
    int* d_rc = nullptr; int64_t* d_rp = nullptr; int32_t* d_col = nullptr; double* d_val = nullptr;
    auto cleanup2 = [&]() { cudaFree(d_rc); cudaFree(d_rp); cudaFree(d_col); cudaFree(d_val); };
    cudaError_t ce;
    std::vector<int> rc(L);
    if ((ce = cudaMalloc(&d_rc, L * 4)) == cudaSuccess) {
      launch_ov_csr(d_inter, d_lp, d_kid, d_anc, AD, d_nt, L, d_rc, nullptr, nullptr, nullptr, nullptr);
      ce = cudaMemcpy(rc.data(), d_rc, L * 4, cudaMemcpyDeviceToHost);
    }
    if (ce == cudaSuccess) {
      row_ptr[0] = 0;
      for (int64_t n = 0; n < L; n++) row_ptr[n + 1] = row_ptr[n] + rc[n];
      if (col && val && row_ptr[L] > 0) {
        if ((ce = cudaMalloc(&d_rp, (L + 1) * 8)) == cudaSuccess && (ce = cudaMalloc(&d_col, row_ptr[L] * 4)) == cudaSuccess &&
            (ce = cudaMalloc(&d_val, row_ptr[L] * 8)) == cudaSuccess && (ce = cudaMemcpy(d_rp, row_ptr, (L + 1) * 8, cudaMemcpyHostToDevice)) == cudaSuccess) {
          launch_ov_csr(d_inter, d_lp, d_kid, d_anc, AD, d_nt, L, nullptr, d_rp, d_col, d_val, nullptr);
          if ((ce = cudaMemcpy(col, d_col, row_ptr[L] * 4, cudaMemcpyDeviceToHost)) == cudaSuccess)
            ce = cudaMemcpy(val, d_val, row_ptr[L] * 8, cudaMemcpyDeviceToHost);
        }
      }
    }
    cleanup2();
    OTRY(ce);
  }
#undef OTRY
  cleanup();
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_overlap(int64_t N, int64_t L, const int64_t* leaf_ptr, const int64_t* leaf_obs,
                                 const int32_t* leaf_kernel_id, const dsmgp_tree* tree, double* D) {
  return overlap_core(N, L, leaf_ptr, leaf_obs, leaf_kernel_id, tree, true, D, nullptr, nullptr, nullptr);
}

// The same matrix in CSR form (row n: the experts m with D[n,m] != 0, ascending).  Call once with col = val = NULL to get
// row_ptr[L+1] (row_ptr[L] = number of non-zeros), then again with col[nnz] (int32) and val[nnz].
extern "C" int32_t dsmgp_overlap_csr(int64_t N, int64_t L, const int64_t* leaf_ptr, const int64_t* leaf_obs,
                                     const int32_t* leaf_kernel_id, const dsmgp_tree* tree, int64_t* row_ptr, int32_t* col, double* val) {
  return overlap_core(N, L, leaf_ptr, leaf_obs, leaf_kernel_id, tree, false, nullptr, row_ptr, col, val);
}

static int32_t chol_host_matrix(double* A, int64_t n, int64_t k, int32_t* info) {
  int32_t rc = standalone_device_check(g_create_error);
  if (rc) return rc;
  const int64_t kp = (k + BLK - 1) / BLK * BLK;          // leading part padded to a block boundary
  const int64_t nn = kp + (n - k);
  LeafMeta m{};
  m.n = (int32_t)nn; m.np = (int32_t)((nn + PAD - 1) / PAD * PAD); m.nb = (m.np + BLK - 1) / BLK; m.nkc = m.np / KC;
  const int np = m.np, nkc = m.nkc;
  const int64_t fd = tiled_doubles(np);
  std::vector<double> P((size_t)fd, 0.0);
  auto map = [&](int64_t i) { return i < k ? i : kp + (i - k); };
  for (int i = 0; i < np; i++) P[tidx(i, i, nkc)] = 1.0;
  for (int64_t c = 0; c < n; c++)
    for (int64_t r = c; r < n; r++) P[tidx((int)map(r), (int)map(c), nkc)] = A[c * n + r];
  std::vector<int4> tasks;
  tasks.push_back(make_int4(0, 0, 0, 0));
  for (int J = 0; J + 1 < m.nb; J++) {
    tasks.push_back(make_int4(0, J + 1, J, 0));
    tasks.push_back(make_int4(0, J + 1, J + 1, 0));
    for (int I = J + 2; I < m.nb; I++) tasks.push_back(make_int4(0, I, J, 0));
  }
  const int64_t nflags = (int64_t)m.nb * (m.nb + 1) / 2;
  double *dF = nullptr, *dW = nullptr, *dWT = nullptr, *dtr = nullptr, *dv = nullptr; LeafMeta* dm = nullptr; LeafScal* ds = nullptr;
  int64_t* doff = nullptr; int* dflags = nullptr; int* dcnt = nullptr; int4* dtasks = nullptr;
  LeafScal sc{}; int gerr = 0;
  const int64_t zero[2] = {0, 0};
  int sms = 148;
  SA_TRY(cudaMalloc(&dF, fd * 8)); SA_TRY(cudaMalloc(&dW, (size_t)m.nb * WBLK_D * 8)); SA_TRY(cudaMalloc(&dWT, (size_t)m.nb * WBLK_D * 8));
  SA_TRY(cudaMalloc(&dtr, 4 * m.nb * 8)); SA_TRY(cudaMalloc(&dv, 2 * (size_t)np * 8)); SA_TRY(cudaMalloc(&dm, sizeof(LeafMeta)));
  SA_TRY(cudaMalloc(&ds, sizeof(LeafScal))); SA_TRY(cudaMalloc(&doff, 16)); SA_TRY(cudaMalloc(&dflags, nflags * 4));
  SA_TRY(cudaMalloc(&dcnt, 64)); SA_TRY(cudaMalloc(&dtasks, tasks.size() * sizeof(int4)));
  SA_TRY(cudaMemcpy(dF, P.data(), fd * 8, cudaMemcpyHostToDevice));
  SA_TRY(cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice));
  SA_TRY(cudaMemset(ds, 0, sizeof(LeafScal))); SA_TRY(cudaMemset(dflags, 0, nflags * 4)); SA_TRY(cudaMemset(dcnt, 0, 64));
  SA_TRY(cudaMemset(dv, 0, 2 * (size_t)np * 8));
  SA_TRY(cudaMemcpy(doff, zero, 16, cudaMemcpyHostToDevice));
  SA_TRY(cudaMemcpy(dtasks, tasks.data(), tasks.size() * sizeof(int4), cudaMemcpyHostToDevice));
  { int cur = 0; cudaGetDevice(&cur); sms = num_sms(cur); }
  {
    Potrf2Args pa{dm, dF, dW, dWT, dv, dv + np, ds, dtr, doff, dtr + 2 * m.nb, dtr + 3 * m.nb, dflags, doff + 1, dtasks,
                  (int)tasks.size(), dcnt, dcnt + 8, (int)(kp / BLK), nullptr};
    launch_potrf2(pa, std::min(sms, (int)tasks.size()), 0);
    SA_TRY(cudaGetLastError());
    SA_TRY(cudaMemcpy(P.data(), dF, fd * 8, cudaMemcpyDeviceToHost));
    SA_TRY(cudaMemcpy(&sc, ds, sizeof(sc), cudaMemcpyDeviceToHost));
    SA_TRY(cudaMemcpy(&gerr, dcnt + 8, sizeof(int), cudaMemcpyDeviceToHost));
  }
  if (gerr != 0) { g_create_error = "device scheduler timeout"; rc = DSMGP_ERR_STATE; goto done; }
  for (int64_t c = 0; c < n; c++)
    for (int64_t r = 0; r < n; r++) A[c * n + r] = (r >= c) ? P[tidx((int)map(r), (int)map(c), nkc)] : 0.0;    // tril!
  if (info) {
    int64_t i = sc.info;                       // 1-based pivot in padded coordinates
    if (i > 0) { i = (i - 1 >= kp) ? (i - 1 - kp) + 1 : i; if (i > n - k) i = 0; }
    *info = (int32_t)i;                        // relative to the trailing block, as LAPACK.potrf!(C) reports it
  }
done:
  cudaFree(dF); cudaFree(dW); cudaFree(dWT); cudaFree(dtr); cudaFree(dv); cudaFree(dm); cudaFree(ds); cudaFree(doff);
  cudaFree(dflags); cudaFree(dcnt); cudaFree(dtasks);
  return rc;
}

extern "C" int32_t dsmgp_potrf(double* A, int64_t n, int32_t* info) {
  if (!A || n <= 0) return DSMGP_ERR_ARG;
  return chol_host_matrix(A, n, 0, info);
}

extern "C" int32_t dsmgp_chol_continue(double* A, int64_t n, int64_t ki, int32_t* info) {
  if (!A || n <= 0 || ki < 1 || ki > n) return DSMGP_ERR_ARG;
  return chol_host_matrix(A, n, ki - 1, info);
}

namespace { struct DelJobH { double* L; int n; const int64_t* rows; int nrows; double* v; }; }

// Row deletion for `count` factors in ONE launch per chunk of 64 deleted rows (one CTA per matrix, one column sweep for all rows
// of a matrix): the device-resident, batched form of what fit.jl:167-195 composes from lowrankupdate!.
extern "C" int32_t dsmgp_chol_delete_rows_batched(int64_t count, const double* const* A, const int64_t* n, const int64_t* const* rows,
                                                  const int64_t* nrows, double* const* out) {
  if (count <= 0 || !A || !n || !rows || !nrows || !out) return DSMGP_ERR_ARG;
  for (int64_t m = 0; m < count; m++) {
    if (!A[m] || n[m] <= 0 || nrows[m] < 0 || nrows[m] >= n[m] || !out[m] || (nrows[m] > 0 && !rows[m])) return DSMGP_ERR_ARG;
    for (int64_t q = 0; q < nrows[m]; q++)
      if (rows[m][q] < 1 || rows[m][q] > n[m] || (q > 0 && rows[m][q] <= rows[m][q - 1])) { g_create_error = "delete_rows: rows must be 1-based ascending"; return DSMGP_ERR_ARG; }
  }
  int32_t rc = standalone_device_check(g_create_error);
  if (rc) return rc;
  constexpr int QMAX = 64;
  std::vector<double*> dL(count, nullptr), dV(count, nullptr); std::vector<int64_t*> dR(count, nullptr);
  DelJobH* djobs = nullptr;
  int64_t maxq = 0;
  for (int64_t m = 0; m < count; m++) maxq = std::max(maxq, nrows[m]);
  for (int64_t m = 0; m < count; m++) {
    SA_TRY(cudaMalloc(&dL[m], n[m] * n[m] * 8));
    SA_TRY(cudaMalloc(&dV[m], (size_t)std::min<int64_t>(std::max<int64_t>(nrows[m], 1), QMAX) * n[m] * 8));
    SA_TRY(cudaMalloc(&dR[m], std::max<int64_t>(nrows[m], 1) * 8));
    SA_TRY(cudaMemcpyAsync(dL[m], A[m], n[m] * n[m] * 8, cudaMemcpyHostToDevice, 0));
    if (nrows[m]) SA_TRY(cudaMemcpyAsync(dR[m], rows[m], nrows[m] * 8, cudaMemcpyHostToDevice, 0));
  }
  SA_TRY(cudaMalloc(&djobs, count * sizeof(DelJobH)));
  for (int64_t q0 = 0; q0 < maxq; q0 += QMAX) {          // later rows are taken from the factor the earlier chunks left behind
    std::vector<DelJobH> jobs(count);
    for (int64_t m = 0; m < count; m++) {
      const int64_t nq = std::max<int64_t>(0, std::min<int64_t>(QMAX, nrows[m] - q0));
      jobs[m] = DelJobH{dL[m], (int)n[m], dR[m] + std::min(q0, nrows[m]), (int)nq, dV[m]};
    }
    SA_TRY(cudaMemcpy(djobs, jobs.data(), count * sizeof(DelJobH), cudaMemcpyHostToDevice));
    launch_delete_rows(djobs, (int)count, 0);
    SA_TRY(cudaGetLastError());
    SA_TRY(cudaDeviceSynchronize());
  }
  for (int64_t m = 0; m < count; m++) {
    std::vector<double> P((size_t)n[m] * n[m]);
    SA_TRY(cudaMemcpy(P.data(), dL[m], n[m] * n[m] * 8, cudaMemcpyDeviceToHost));
    std::vector<int64_t> keep;
    for (int64_t i = 0, q = 0; i < n[m]; i++) { if (q < nrows[m] && rows[m][q] - 1 == i) { q++; continue; } keep.push_back(i); }
    const int64_t k = (int64_t)keep.size();
    for (int64_t c = 0; c < k; c++)
      for (int64_t r = 0; r < k; r++) out[m][c * k + r] = (r >= c) ? P[keep[c] * n[m] + keep[r]] : 0.0;
  }
done:
  for (int64_t m = 0; m < count; m++) { cudaFree(dL[m]); cudaFree(dV[m]); cudaFree(dR[m]); }
  cudaFree(djobs);
  return rc;
}

extern "C" int32_t dsmgp_chol_delete_rows(const double* A, int64_t n, const int64_t* rows, int64_t nrows, double* out) {
  if (!A || n <= 0 || !rows || nrows < 0 || nrows >= n || !out) return DSMGP_ERR_ARG;
  return dsmgp_chol_delete_rows_batched(1, &A, &n, &rows, &nrows, &out);
}
