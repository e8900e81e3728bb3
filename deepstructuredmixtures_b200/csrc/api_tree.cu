// libdsmgp.so : region-graph construction primitives on the device (kernels in k_tree.cu).
// The random draws and the recursion of treeStructure.jl:23-243 stay on the host (bit-identical partitions); the passes over the
// data (sorting a node's column, per-dimension ranges, the K-way stable partition of _buildSplit) run on index lists in HBM.
#include "handle.h"

using namespace dsm;
#define g_create_error (dsm::create_error())

namespace dsm {
void part_iota(int* a, int64_t n, cudaStream_t st);
void part_gather(const double* x, int64_t N, int d, const int* idx, int64_t n, double* out, cudaStream_t st);
int part_range_blocks(int64_t n);
void part_range(const double* x, int64_t N, int D, const int* idx, int64_t n, double* partial, cudaStream_t st);
void part_count(const double* x, int64_t N, int d, const int* idx, int64_t n, const double* lower, const double* upper, int K, int* counts, cudaStream_t st);
void part_scatter(const double* x, int64_t N, int d, const int* idx, int64_t n, const double* lower, const double* upper, int K,
                  const int64_t* base, const int64_t* child_off, int* arena, cudaStream_t st);
size_t part_sort_temp_bytes(int64_t n);
cudaError_t part_sort(void* temp, size_t temp_bytes, const double* in, double* out, int64_t n, cudaStream_t st);
}

struct dsmgp_partition {
  int64_t N = 0, D = 0;
  int device = 0;
  double* d_x = nullptr;
  int* arena = nullptr; int64_t arena_cap = 0, arena_used = 0;     // index lists (0-based global rows), one segment per node
  std::vector<int64_t> off, size;                                  // per node
  double *d_a = nullptr, *d_b = nullptr; int64_t ab_cap = 0;       // gather / sort buffers
  void* d_temp = nullptr; size_t temp_cap = 0;
  int* d_counts = nullptr; int64_t* d_base = nullptr; int64_t cb_cap = 0;
  double* d_bounds = nullptr; int64_t* d_child_off = nullptr;
  double* d_partial = nullptr;
  std::string err;
  ~dsmgp_partition() {
    cudaFree(d_x); cudaFree(arena); cudaFree(d_a); cudaFree(d_b); cudaFree(d_temp); cudaFree(d_counts); cudaFree(d_base);
    cudaFree(d_bounds); cudaFree(d_child_off); cudaFree(d_partial);
  }
};

#define PTRY(p, expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { (p)->err = std::string(#expr) + ": " + cudaGetErrorString(e_); \
    g_create_error = (p)->err; return e_ == cudaErrorMemoryAllocation ? DSMGP_ERR_OOM : DSMGP_ERR_CUDA; } } while (0)

static int32_t part_reserve(dsmgp_partition* p, int64_t extra) {
  if (p->arena_used + extra <= p->arena_cap) return DSMGP_OK;
  int64_t cap = std::max<int64_t>(p->arena_cap * 2, p->arena_used + extra);
  int* na = nullptr;
  PTRY(p, cudaMalloc(&na, cap * sizeof(int)));
  if (p->arena_used) PTRY(p, cudaMemcpy(na, p->arena, p->arena_used * sizeof(int), cudaMemcpyDeviceToDevice));
  cudaFree(p->arena);
  p->arena = na; p->arena_cap = cap;
  return DSMGP_OK;
}

extern "C" int32_t dsmgp_part_create(const double* x, int64_t N, int64_t D, dsmgp_partition** out) {
  if (out) *out = nullptr;
  if (!x || N <= 0 || D <= 0 || !out || N >= (int64_t(1) << 31)) { g_create_error = "part_create: bad argument"; return DSMGP_ERR_ARG; }
  int32_t rc = standalone_device_check(g_create_error);
  if (rc) return rc;
  dsmgp_partition* p = new dsmgp_partition();
  p->N = N; p->D = D;
  cudaGetDevice(&p->device);
  auto fail = [&](int32_t code) { delete p; return code; };
  if (cudaMalloc(&p->d_x, (size_t)N * D * sizeof(double)) != cudaSuccess) { g_create_error = "part_create: out of device memory"; return fail(DSMGP_ERR_OOM); }
  if (cudaMemcpy(p->d_x, x, (size_t)N * D * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) { g_create_error = "part_create: upload failed"; return fail(DSMGP_ERR_CUDA); }
  if ((rc = part_reserve(p, 4 * N))) return fail(rc);
  part_iota(p->arena, N, 0);
  p->off.push_back(0); p->size.push_back(N); p->arena_used = N;
  p->ab_cap = N;
  p->temp_cap = part_sort_temp_bytes(N);
  if (cudaMalloc(&p->d_a, N * sizeof(double)) || cudaMalloc(&p->d_b, N * sizeof(double)) || cudaMalloc(&p->d_temp, std::max<size_t>(p->temp_cap, 16)) ||
      cudaMalloc(&p->d_bounds, 64 * sizeof(double)) || cudaMalloc(&p->d_child_off, 32 * sizeof(int64_t)) ||
      cudaMalloc(&p->d_partial, (size_t)part_range_blocks(N) * D * 2 * sizeof(double))) { g_create_error = "part_create: out of device memory"; return fail(DSMGP_ERR_OOM); }
  p->cb_cap = ((N + 255) / 256) * 32;
  if (cudaMalloc(&p->d_counts, p->cb_cap * sizeof(int)) || cudaMalloc(&p->d_base, p->cb_cap * sizeof(int64_t))) { g_create_error = "part_create: out of device memory"; return fail(DSMGP_ERR_OOM); }
  if (cudaDeviceSynchronize() != cudaSuccess) { g_create_error = "part_create: device error"; return fail(DSMGP_ERR_CUDA); }
  *out = p;
  return DSMGP_OK;
}

extern "C" void dsmgp_part_destroy(dsmgp_partition* p) { if (p) { cudaSetDevice(p->device); delete p; } }

extern "C" int64_t dsmgp_part_size(const dsmgp_partition* p, int64_t node) {
  if (!p || node < 0 || node >= (int64_t)p->size.size()) return -1;
  return p->size[node];
}

#define NODE_CHECK(p, node) if (!(p) || (node) < 0 || (node) >= (int64_t)(p)->size.size()) { g_create_error = "partition: bad node"; return DSMGP_ERR_ARG; }

// X.max / X.min per dimension over the node's rows (_buildSum, treeStructure.jl:233)
extern "C" int32_t dsmgp_part_range(dsmgp_partition* p, int64_t node, double* mins, double* maxs) {
  NODE_CHECK(p, node);
  if (!mins || !maxs) return DSMGP_ERR_ARG;
  cudaSetDevice(p->device);
  const int64_t n = p->size[node];
  const int nb = part_range_blocks(n);
  for (int64_t d = 0; d < p->D; d++) { mins[d] = std::numeric_limits<double>::infinity(); maxs[d] = -std::numeric_limits<double>::infinity(); }
  if (n == 0) return DSMGP_OK;
  part_range(p->d_x, p->N, (int)p->D, p->arena + p->off[node], n, p->d_partial, 0);
  std::vector<double> part((size_t)nb * p->D * 2);
  PTRY(p, cudaMemcpy(part.data(), p->d_partial, part.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (int b = 0; b < nb; b++)
    for (int64_t d = 0; d < p->D; d++) { mins[d] = std::min(mins[d], part[(b * p->D + d) * 2]); maxs[d] = std::max(maxs[d], part[(b * p->D + d) * 2 + 1]); }
  return DSMGP_OK;
}

// the node's column d in ascending order: every query of getSplits (treeStructure.jl:23-129) becomes a binary search
extern "C" int32_t dsmgp_part_sorted_column(dsmgp_partition* p, int64_t node, int64_t d, double* sorted) {
  NODE_CHECK(p, node);
  if (d < 0 || d >= p->D || !sorted) return DSMGP_ERR_ARG;
  cudaSetDevice(p->device);
  const int64_t n = p->size[node];
  if (n == 0) return DSMGP_OK;
  part_gather(p->d_x, p->N, (int)d, p->arena + p->off[node], n, p->d_a, 0);
  PTRY(p, part_sort(p->d_temp, p->temp_cap, p->d_a, p->d_b, n, 0));
  PTRY(p, cudaMemcpy(sorted, p->d_b, n * sizeof(double), cudaMemcpyDeviceToHost));
  return DSMGP_OK;
}

// children of _buildSplit (treeStructure.jl:176-199): child k = rows with lower[k] < x_d <= upper[k], in the node's own order
extern "C" int32_t dsmgp_part_split(dsmgp_partition* p, int64_t node, int64_t d, const double* lower, const double* upper, int64_t K,
                                    int64_t* children, int64_t* sizes) {
  NODE_CHECK(p, node);
  if (d < 0 || d >= p->D || !lower || !upper || K <= 0 || K > 32 || !children || !sizes) { g_create_error = "part_split: bad argument"; return DSMGP_ERR_ARG; }
  cudaSetDevice(p->device);
  const int64_t n = p->size[node];
  const int64_t nb = (n + 255) / 256;
  std::vector<double> bounds(64);
  for (int64_t k = 0; k < K; k++) { bounds[k] = lower[k]; bounds[32 + k] = upper[k]; }
  PTRY(p, cudaMemcpy(p->d_bounds, bounds.data(), 64 * sizeof(double), cudaMemcpyHostToDevice));
  std::vector<int> counts((size_t)nb * K, 0);
  if (n > 0) {
    part_count(p->d_x, p->N, (int)d, p->arena + p->off[node], n, p->d_bounds, p->d_bounds + 32, (int)K, p->d_counts, 0);
    PTRY(p, cudaMemcpy(counts.data(), p->d_counts, counts.size() * sizeof(int), cudaMemcpyDeviceToHost));
  }
  std::vector<int64_t> base((size_t)nb * K, 0), tot(K, 0), coff(32, 0);
  for (int64_t b = 0; b < nb; b++) for (int64_t k = 0; k < K; k++) { base[b * K + k] = tot[k]; tot[k] += counts[b * K + k]; }
  int64_t need = 0;
  for (int64_t k = 0; k < K; k++) need += tot[k];
  { int32_t rc = part_reserve(p, need); if (rc) return rc; }
  for (int64_t k = 0; k < K; k++) {
    coff[k] = p->arena_used;
    children[k] = (int64_t)p->size.size(); sizes[k] = tot[k];
    p->off.push_back(p->arena_used); p->size.push_back(tot[k]);
    p->arena_used += tot[k];
  }
  if (n > 0 && need > 0) {
    PTRY(p, cudaMemcpy(p->d_base, base.data(), base.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    PTRY(p, cudaMemcpy(p->d_child_off, coff.data(), 32 * sizeof(int64_t), cudaMemcpyHostToDevice));
    part_scatter(p->d_x, p->N, (int)d, p->arena + p->off[node], n, p->d_bounds, p->d_bounds + 32, (int)K, p->d_base, p->d_child_off, p->arena, 0);
    PTRY(p, cudaGetLastError());
  }
  return DSMGP_OK;
}

// GPNode.obs of a leaf (treeStructure.jl:245-307): the node's rows, 1-based, ascending
extern "C" int32_t dsmgp_part_rows(dsmgp_partition* p, int64_t node, int64_t* obs) {
  NODE_CHECK(p, node);
  if (!obs) return DSMGP_ERR_ARG;
  cudaSetDevice(p->device);
  const int64_t n = p->size[node];
  std::vector<int> tmp(n);
  if (n) PTRY(p, cudaMemcpy(tmp.data(), p->arena + p->off[node], n * sizeof(int), cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < n; i++) obs[i] = (int64_t)tmp[i] + 1;
  return DSMGP_OK;
}
