// Launchers of the INT8-tensor-core block products (ozaki.cuh).
#include "ozaki.cuh"

namespace dsm {

cudaError_t oz_init_kernels() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(oz::gemm_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::Cfg<7>::SMEM))) return e;
  if ((e = cudaFuncSetAttribute(oz::gemm_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::Cfg<8>::SMEM))) return e;
  return cudaSuccess;
}

// 2-D tensor map over the slice pool seen as rows of 128-byte core matrices; box = the slice tiles one round needs of a k-step
// (32 rows = 4096 bytes per slice tile; at most 8 tiles = 256 rows, the largest box a tensor map takes).
// No swizzle, no interleave: a box lands in shared memory exactly as it lies in global memory, which is the UMMA
// no-swizzle K-major layout the slicing kernel wrote.  cuTensorMapEncodeTiled comes from the driver through the runtime
// (the library does not link libcuda).
int oz_make_map(void* map128, const void* pool, size_t bytes, int box_rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess) { fn = nullptr; return 1; }
  }
  static_assert(sizeof(CUtensorMap) == 128, "tensor map size");
  cuuint64_t dims[2] = {128, (cuuint64_t)(bytes / 128)};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {128, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(reinterpret_cast<CUtensorMap*>(map128), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(pool), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

void launch_oz_slice(int S, const OzJob* jobs, int njobs, int pass, unsigned long long* rowmax, double* scale, int8_t* pool, cudaStream_t st) {
  if (njobs <= 0) return;
  if (S == 7) oz::slice_kernel<7><<<njobs, 256, 0, st>>>(jobs, pass, rowmax, scale, pool);
  else oz::slice_kernel<8><<<njobs, 256, 0, st>>>(jobs, pass, rowmax, scale, pool);
}

void launch_oz_gemm(int S, const void* maps256, const OzTile* tiles, int ntiles, const double* scale, int nctas, cudaStream_t st, long long* trace) {
  if (ntiles <= 0) return;
  CUtensorMap map0, map1;       // round 0 / round 1 boxes
  memcpy(&map0, maps256, sizeof(map0));
  memcpy(&map1, static_cast<const unsigned char*>(maps256) + 128, sizeof(map1));
  const int grid = ntiles < nctas ? ntiles : nctas;
  if (S == 7) oz::gemm_kernel<7><<<grid, oz::OZ_THREADS, oz::Cfg<7>::SMEM, st>>>(map0, map1, tiles, ntiles, scale, trace);
  else oz::gemm_kernel<8><<<grid, oz::OZ_THREADS, oz::Cfg<8>::SMEM, st>>>(map0, map1, tiles, ntiles, scale, trace);
}

int oz_round_slices(int S, int r) { return S == 7 ? oz::Cfg<7>::nsl(r) : oz::Cfg<8>::nsl(r); }

void launch_oz_parts(const OzPartArgs& a, int nparts, cudaStream_t st) {
  if (nparts <= 0) return;
  oz::parts_kernel<<<nparts, BLK, 0, st>>>(a);
}

void launch_oz_setflags(const OzPart* parts, int n, const int64_t* flag_off, int* flags, cudaStream_t st) {
  if (n > 0) oz::setflags_kernel<<<(n + 255) / 256, 256, 0, st>>>(parts, n, flag_off, flags);
}

}  // namespace dsm
