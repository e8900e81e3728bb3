// Region-graph construction primitives on the device: the data-dependent part of treeStructure.jl:23-243.
//
// The reference builds the region graph with repeated `findall` / `median` / `sum(. <= s)` passes over N-vectors per node
// (treeStructure.jl:40,49,56-57,148,181) and copies X[idx,:] at every level.  Here X stays on the device, every node of the
// construction is an index list in a device arena, and the host recursion (which keeps every random draw: Beta, rand(1:2),
// Categorical -- partitions must stay bit-identical to the host builder) asks for exactly three things:
//   * the sorted column d of a node      -> every query of getSplits (min, max, median of a sub-range, counts) is a binary search
//   * the per-dimension range of a node  -> _buildSum's phi = max - min (treeStructure.jl:233-235)
//   * a stable K-way partition by (lower, upper] intervals -> the children of _buildSplit (treeStructure.jl:176-199)
// Sorting uses CUB's device radix sort (NVIDIA's primitive library, like cuBLAS for a plain GEMM); the gather, range and
// partition kernels are written here.
#include <cub/device/device_radix_sort.cuh>

#include "args.h"

namespace dsm {

__global__ void part_gather_kernel(const double* x, int64_t N, int d, const int* idx, int64_t n, double* out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = x[(int64_t)d * N + idx[i]];
}

// per-block min / max of every dimension over the rows idx[0..n): partial[(b * D + d) * 2 + {0,1}]
__global__ void __launch_bounds__(256) part_range_kernel(const double* x, int64_t N, int D, const int* idx, int64_t n, double* partial) {
  __shared__ double smin[8], smax[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int d = 0; d < D; d++) {
    double mn = INFINITY, mx = -INFINITY;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + tid; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const double v = x[(int64_t)d * N + idx[i]];
      mn = fmin(mn, v); mx = fmax(mx, v);
    }
    for (int o = 16; o > 0; o >>= 1) { mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if (lane == 0) { smin[warp] = mn; smax[warp] = mx; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; w++) { mn = fmin(mn, smin[w]); mx = fmax(mx, smax[w]); }
      partial[((int64_t)blockIdx.x * D + d) * 2] = mn; partial[((int64_t)blockIdx.x * D + d) * 2 + 1] = mx;
    }
    __syncthreads();
  }
}

// interval of a value: k with lower[k] < v <= upper[k] (first match), or -1
__device__ __forceinline__ int part_class(double v, const double* lower, const double* upper, int K) {
  for (int k = 0; k < K; k++) if (v > lower[k] && v <= upper[k]) return k;
  return -1;
}

// pass 1: counts[(b * K) + k] = elements of block b (256 consecutive list entries) in interval k
__global__ void __launch_bounds__(256) part_count_kernel(const double* x, int64_t N, int d, const int* idx, int64_t n,
                                                         const double* lower, const double* upper, int K, int* counts) {
  __shared__ int sc[32];
  const int tid = threadIdx.x;
  if (tid < K) sc[tid] = 0;
  __syncthreads();
  const int64_t i = blockIdx.x * 256ll + tid;
  if (i < n) { const int k = part_class(x[(int64_t)d * N + idx[i]], lower, upper, K); if (k >= 0) atomicAdd(&sc[k], 1); }
  __syncthreads();
  if (tid < K) counts[(int64_t)blockIdx.x * K + tid] = sc[tid];
}

// pass 2: stable scatter.  base[(b * K) + k] = first output position (inside child k's list) of block b's elements.
__global__ void __launch_bounds__(256) part_scatter_kernel(const double* x, int64_t N, int d, const int* idx, int64_t n,
                                                           const double* lower, const double* upper, int K, const int64_t* base,
                                                           const int64_t* child_off, int* arena) {
  __shared__ int wcnt[8][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t i = blockIdx.x * 256ll + tid;
  int row = 0, k = -1;
  if (i < n) { row = idx[i]; k = part_class(x[(int64_t)d * N + row], lower, upper, K); }
  int rank = 0;
  for (int q = 0; q < K; q++) {
    const unsigned m = __ballot_sync(0xffffffffu, k == q);
    if (k == q) rank = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) wcnt[warp][q] = __popc(m);
  }
  __syncthreads();
  if (k >= 0) {
    int pre = 0;
    for (int w = 0; w < warp; w++) pre += wcnt[w][k];
    arena[child_off[k] + base[(int64_t)blockIdx.x * K + k] + pre + rank] = row;
  }
}

__global__ void part_iota_kernel(int* a, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] = (int)i;
}

// ---- host-callable wrappers (api_tree.cu) ---------------------------------------------------------------------------------
void part_iota(int* a, int64_t n, cudaStream_t st) { part_iota_kernel<<<1184, 256, 0, st>>>(a, n); }
void part_gather(const double* x, int64_t N, int d, const int* idx, int64_t n, double* out, cudaStream_t st) {
  part_gather_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 4736), 256, 0, st>>>(x, N, d, idx, n, out);
}
int part_range_blocks(int64_t n) { return (int)std::min<int64_t>((n + 2047) / 2048, 592); }
void part_range(const double* x, int64_t N, int D, const int* idx, int64_t n, double* partial, cudaStream_t st) {
  part_range_kernel<<<part_range_blocks(n), 256, 0, st>>>(x, N, D, idx, n, partial);
}
void part_count(const double* x, int64_t N, int d, const int* idx, int64_t n, const double* lower, const double* upper, int K, int* counts, cudaStream_t st) {
  part_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, N, d, idx, n, lower, upper, K, counts);
}
void part_scatter(const double* x, int64_t N, int d, const int* idx, int64_t n, const double* lower, const double* upper, int K,
                  const int64_t* base, const int64_t* child_off, int* arena, cudaStream_t st) {
  part_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, N, d, idx, n, lower, upper, K, base, child_off, arena);
}
size_t part_sort_temp_bytes(int64_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, bytes, (const double*)nullptr, (double*)nullptr, (int)n);
  return bytes;
}
cudaError_t part_sort(void* temp, size_t temp_bytes, const double* in, double* out, int64_t n, cudaStream_t st) {
  return cub::DeviceRadixSort::SortKeys(temp, temp_bytes, in, out, (int)n, 0, 64, st);
}

}  // namespace dsm
