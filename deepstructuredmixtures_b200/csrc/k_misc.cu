#include "args.h"
#include "route.cuh"
#include "tree.cuh"
namespace dsm {
void launch_route(const RouteArgs& a, bool fill, cudaStream_t st) {
  const unsigned nb = (unsigned)((a.T + 255) / 256);
  if (fill) route_kernel<true><<<nb, 256, 0, st>>>(a); else route_kernel<false><<<nb, 256, 0, st>>>(a);
}
void launch_tree_eval(const TreeEvalArgs& a, cudaStream_t st) { tree_eval_kernel<<<1, 256, 0, st>>>(a); }
void launch_derive(const DeriveArgs& a, cudaStream_t st) { if (a.nslots > 0) derive_kernel<<<(a.nslots + 127) / 128, 128, 0, st>>>(a); }
void launch_opt_step(const OptArgs& a, cudaStream_t st) { opt_step_kernel<<<1, (a.H + 63) / 64 * 64, 0, st>>>(a); }
void launch_mix(const MixArgs& a, cudaStream_t st) { mix_kernel<<<(unsigned)((a.T + 127) / 128), 128, 0, st>>>(a); }
}  // namespace dsm
namespace dsm {
// Row / column deletion from lower Cholesky factors (the operation fit.jl:167-195 composes from lowrankupdate!,
// AdvancedCholeskey.jl:20-59, implemented correctly: SURVEY App. B Q7).  Deleting row i turns the trailing block into the factor of
// L33 L33' + u u' (u = L[i+1:, i]): a Givens rank-1 UPDATE that sweeps the columns behind i.  The reference runs one full sweep per
// deleted row; here ONE sweep over the columns serves every deleted row of a matrix: at column k the rotations of all update
// vectors that are already active are applied back to back while the column is in registers (the arithmetic and its order per
// element are exactly those of the row-by-row sweeps, so the result is bit-identical), and when k is itself a deleted row its
// freshly updated column becomes the next update vector.  Column traffic and barriers drop by the number of deleted rows.
// One CTA per matrix: a batch of matrices is one launch (dsmgp_chol_delete_rows_batched).
struct DelJob { double* L; int n; const int64_t* rows; int nrows; double* v; };      // v: [nrows][n] scratch
constexpr int DEL_QMAX = 64;
__global__ void __launch_bounds__(NTHREADS) delete_rows_kernel(const DelJob* jobs) {
  __shared__ double cs[2 * DEL_QMAX];
  const DelJob jb = jobs[blockIdx.x];
  double* Lf = jb.L; const int n = jb.n; const int tid = threadIdx.x;
  if (jb.nrows <= 0) return;
  int nact = 0, next = 0;
  for (int k = (int)jb.rows[0] - 1; k < n; k++) {
    if (nact > 0) {
      if (tid == 0) {
        double f = Lf[(int64_t)k * n + k];
        for (int j = 0; j < nact; j++) {
          const double g = jb.v[(int64_t)j * n + k];
          double c, s, r;
          if (g == 0.0) { c = 1.0; s = 0.0; r = f; }
          else if (f == 0.0) { c = 0.0; s = 1.0; r = g; }
          else { r = hypot(f, g); if (fabs(f) > fabs(g) && f < 0) r = -r; c = f / r; s = g / r; }
          cs[2 * j] = c; cs[2 * j + 1] = s; f = r;
        }
        Lf[(int64_t)k * n + k] = f;
      }
      __syncthreads();
      for (int r = k + 1 + tid; r < n; r += NTHREADS) {
        double a = Lf[(int64_t)k * n + r];
        for (int j = 0; j < nact; j++) {
          const double c = cs[2 * j], s = cs[2 * j + 1];
          const double b = jb.v[(int64_t)j * n + r];
          const double a2 = c * a + s * b;
          jb.v[(int64_t)j * n + r] = -s * a + c * b;
          a = a2;
        }
        Lf[(int64_t)k * n + r] = a;
      }
      __syncthreads();
    }
    if (next < jb.nrows && (int)jb.rows[next] - 1 == k) {          // column k (now final) is the update vector of its own deletion
      for (int r = k + 1 + tid; r < n; r += NTHREADS) jb.v[(int64_t)nact * n + r] = Lf[(int64_t)k * n + r];
      nact++; next++;
      __syncthreads();
    }
  }
}

// tiled factor -> dense column-major n x n (lower triangle; upper zero)
__global__ void untile_kernel(const double* Ft, int nkc, int n, double* out) {
  const int c = blockIdx.x;
  for (int r = threadIdx.x; r < n; r += blockDim.x) out[(int64_t)c * n + r] = (r >= c) ? Ft[tidx(r, c, nkc)] : 0.0;
}
void launch_untile(const double* Ft, int nkc, int n, double* out, cudaStream_t st) { untile_kernel<<<n, 256, 0, st>>>(Ft, nkc, n, out); }
void launch_delete_rows(const void* jobs_dev, int njobs, cudaStream_t st) {
  if (njobs > 0) delete_rows_kernel<<<njobs, NTHREADS, 0, st>>>(static_cast<const DelJob*>(jobs_dev));
}

// ---- shared Cholesky of fit! (fit.jl:132-143, 208-292): block rows I < jb of a SHARE_PREFIX expert are the source expert's
// tiles (I, 0 .. 8(I+1)-1) verbatim (same observations, same hyper-parameters => bit-identical factor).  grid = (experts, max jb).
__global__ void __launch_bounds__(NTHREADS) share_copy_kernel(const LeafMeta* meta, const int4* share, const int* slots, double* F) {
  const int s = slots[blockIdx.x];
  const int4 sh = share[s];
  const int I = blockIdx.y;
  if (sh.x != SHARE_PREFIX || I >= sh.z) return;
  const LeafMeta md = meta[s], ms = meta[sh.y];
  const double2* src = reinterpret_cast<const double2*>(F + ms.foff + tile_off(I, 0, ms.nkc));
  double2* dst = reinterpret_cast<double2*>(F + md.foff + tile_off(I, 0, md.nkc));
  const int n2 = (I + 1) * (BLK / KC) * TILE_D / 2;
  for (int i = threadIdx.x; i < n2; i += NTHREADS) dst[i] = src[i];
}
void launch_share_copy(const LeafMeta* meta, const int4* share, const int* slots, int nslots, int max_jb, double* F, cudaStream_t st) {
  if (nslots <= 0 || max_jb <= 0) return;
  share_copy_kernel<<<dim3(nslots, max_jb), NTHREADS, 0, st>>>(meta, share, slots, F);
}
// rows (and the potrf info) of SHARE_ALIAS experts = those of their source (fit.jl:132-143 copies factors and alpha)
__global__ void rows_alias_kernel(const LeafMeta* meta, const int4* share, int nslots, double* rows, int row_width, LeafScal* scal) {
  for (int s = blockIdx.x; s < nslots; s += gridDim.x) {
    const int4 sh = share[s];
    if (sh.x != SHARE_ALIAS) continue;
    const int ld = meta[s].leaf, ls = meta[sh.y].leaf;
    for (int k = threadIdx.x; k < row_width; k += blockDim.x) rows[(int64_t)ld * row_width + k] = rows[(int64_t)ls * row_width + k];
    if (threadIdx.x == 0) scal[s] = scal[sh.y];
  }
}
void launch_rows_alias(const LeafMeta* meta, const int4* share, int nslots, double* rows, int row_width, LeafScal* scal, cudaStream_t st) {
  if (nslots <= 0) return;
  rows_alias_kernel<<<nslots < 1184 ? nslots : 1184, 32, 0, st>>>(meta, share, nslots, rows, row_width, scal);
}

// ---- overlap matrix of getOverlap (fit.jl:12-39) -------------------------------------------------------------
// D[n,m] = 1 - |obs_n \ obs_m| / |obs_n| = f(|obs_n ^ obs_m|) for experts whose lowest common ancestor is a sum node.
// The reference xors N-bit sets for every pair (O(L^2 N)); here every POINT enumerates the pairs of experts that
// contain it (it belongs to one expert per sum-node branch), so the work is sum_p k_p^2 integer atomics.
__global__ void ov_count_kernel(const int64_t* obs, int64_t total, int* cntp) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(cntp + (obs[i] - 1), 1);
}
__global__ void ov_fill_kernel(const int64_t* obs, const int64_t* leaf_ptr, int L, const int64_t* poff, int* fill, int* plist) {
  const int l = blockIdx.x;
  for (int64_t i = leaf_ptr[l] + threadIdx.x; i < leaf_ptr[l + 1]; i += blockDim.x) {
    const int64_t p = obs[i] - 1;
    plist[poff[p] + atomicAdd(fill + p, 1)] = l;
  }
  (void)L;
}
// one warp per point: every unordered pair of its experts gets one count (both orientations)
__global__ void ov_pairs_kernel(const int64_t* poff, const int* plist, int64_t N, int64_t L, int* inter) {
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (int64_t p = warp; p < N; p += nw) {
    const int64_t o = poff[p];
    const int k = (int)(poff[p + 1] - o);
    const int npairs = k * (k - 1) / 2;
    for (int q = lane; q < npairs; q += 32) {
      int i = (int)((sqrtf(8.0f * q + 1.0f) + 1.0f) * 0.5f);      // q = i(i-1)/2 + j, j < i
      while (i * (i - 1) / 2 > q) i--;
      while ((i + 1) * i / 2 <= q) i++;
      const int j = q - i * (i - 1) / 2;
      const int64_t a = plist[o + i], b = plist[o + j];
      atomicAdd(inter + a + b * L, 1);
      atomicAdd(inter + b + a * L, 1);
    }
  }
}
// anc: [L][AD] node ids from the root down to the expert (-1 padded).  D[n,m] of getOverlap (fit.jl:12-39).
__device__ __forceinline__ double ov_value(const int* inter, const int64_t* leaf_ptr, const int* kid, const int* anc, int AD,
                                           const int* node_type, int64_t L, int64_t n, int64_t m) {
  if (n == m) return 0.0;
  int lca = -1;
  for (int d = 0; d < AD; d++) {
    const int an = anc[n * AD + d], am = anc[m * AD + d];
    if (an < 0 || an != am) break;
    lca = an;
  }
  if (lca < 0 || node_type[lca] < 2) return 0.0;                   // DSMGP_NODE_SUM / DSMGP_NODE_KSUM
  const int64_t cn = leaf_ptr[n + 1] - leaf_ptr[n];
  const int64_t dn = (kid[n] == kid[m]) ? cn - (int64_t)inter[n + m * L] : 0;      // sum(xor & obs_n) * (kernelid equal)
  return 1.0 - (double)dn / (double)cn;                            // fit.jl:31
}
// D is L x L column-major.
__global__ void ov_finish_kernel(const int* inter, const int64_t* leaf_ptr, const int* kid, const int* anc, int AD,
                                 const int* node_type, int64_t L, double* D) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < L * L; e += (int64_t)gridDim.x * blockDim.x)
    D[e] = ov_value(inter, leaf_ptr, kid, anc, AD, node_type, L, e % L, e / L);
}
// Sparse form (CSR by row n, columns ascending): one warp per row.  col == nullptr: count pass (row_cnt[n] = non-zeros of row n).
__global__ void ov_csr_kernel(const int* inter, const int64_t* leaf_ptr, const int* kid, const int* anc, int AD, const int* node_type,
                              int64_t L, int* row_cnt, const int64_t* row_ptr, int32_t* col, double* val) {
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (int64_t n = warp; n < L; n += nw) {
    int64_t pos = col != nullptr ? row_ptr[n] : 0;
    int cnt = 0;
    for (int64_t m0 = 0; m0 < L; m0 += 32) {
      const int64_t m = m0 + lane;
      const double v = m < L ? ov_value(inter, leaf_ptr, kid, anc, AD, node_type, L, n, m) : 0.0;
      const unsigned mask = __ballot_sync(0xffffffffu, v != 0.0);
      if (col != nullptr && v != 0.0) { const int r = __popc(mask & ((1u << lane) - 1u)); col[pos + r] = (int32_t)m; val[pos + r] = v; }
      pos += __popc(mask); cnt += __popc(mask);
    }
    if (col == nullptr && lane == 0) row_cnt[n] = cnt;
  }
}
void launch_ov_count(const int64_t* obs, int64_t total, int* cntp, cudaStream_t st) { ov_count_kernel<<<1184, 256, 0, st>>>(obs, total, cntp); }
void launch_ov_fill(const int64_t* obs, const int64_t* leaf_ptr, int L, const int64_t* poff, int* fill, int* plist, cudaStream_t st) {
  ov_fill_kernel<<<L, 256, 0, st>>>(obs, leaf_ptr, L, poff, fill, plist);
}
void launch_ov_pairs(const int64_t* poff, const int* plist, int64_t N, int64_t L, int* inter, cudaStream_t st) {
  ov_pairs_kernel<<<1184, 256, 0, st>>>(poff, plist, N, L, inter);
}
void launch_ov_csr(const int* inter, const int64_t* leaf_ptr, const int* kid, const int* anc, int AD, const int* node_type, int64_t L,
                   int* row_cnt, const int64_t* row_ptr, int32_t* col, double* val, cudaStream_t st) {
  ov_csr_kernel<<<1184, 256, 0, st>>>(inter, leaf_ptr, kid, anc, AD, node_type, L, row_cnt, row_ptr, col, val);
}
void launch_ov_finish(const int* inter, const int64_t* leaf_ptr, const int* kid, const int* anc, int AD, const int* node_type,
                      int64_t L, double* D, cudaStream_t st) {
  ov_finish_kernel<<<1184, 256, 0, st>>>(inter, leaf_ptr, kid, anc, AD, node_type, L, D);
}
}  // namespace dsm
