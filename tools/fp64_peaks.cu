// FP64 roofline probes for B200 (sm_100a): the denominators bench.py reports against.
// MEASURED_PEAKS.json (driver-written) has only HBM and bf16 numbers; this path is FP64.
//   dfma      : register-resident DFMA chains (CUDA-core FP64 roof)
//   dmma      : DMMA.8x8x4 chains (mma.sync m8n8k4 f64; the only native FP64 MMA shape on sm_100a)
//   dmma_lat  : one dependent DMMA chain per warp (latency)
//   dgemm     : cuBLAS DGEMM n^3 (library roof for a dense FP64 contraction)
//   exp       : FP64 exp() throughput (binds the ARD Gram build)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu -lcublas
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int CH>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double acc[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters) {
  double c0[CH], c1[CH];
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
#pragma unroll
  for (int i = 0; i < CH; i++) { c0[i] = i; c1[i] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) exp_kernel(double* out, int iters, double step) {
  double x = -1e-3 * threadIdx.x, s = 0;
  for (int it = 0; it < iters; it++) {
    s += exp(x); x -= step;
    s += exp(x * 0.5); s += exp(x * 0.25); s += exp(x * 0.125);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F f, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  std::vector<float> ts;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ts.push_back(ms);
  }
  return *std::min_element(ts.begin(), ts.end());
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * 148 * 64 * 256));

  for (int wps = 4; wps <= 32; wps *= 2) {  // warps per SM
    int blocks = sms * (wps * 32 / 256 > 0 ? wps * 32 / 256 : 1);
    int threads = wps * 32 >= 256 ? 256 : wps * 32;
    int iters = 20000;
    double ms = time_ms([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters, 0.999, 1e-3); }, 5);
    double fl = 2.0 * 8 * iters * (double)blocks * threads;
    printf(", \"dfma_tflops_w%d\": %.2f", wps, fl / ms * 1e-9);
    ms = time_ms([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters); }, 5);
    fl = 2.0 * 256 * 8 * iters * (double)blocks * threads / 32;
    printf(", \"dmma_tflops_w%d\": %.2f", wps, fl / ms * 1e-9);
  }
  {  // 2 independent chains: latency-ish view (chains=1 approximated by CH=1)
    int blocks = sms, threads = 128, iters = 20000;
    double ms = time_ms([&] { dmma_kernel<1><<<blocks, threads>>>(out, iters); }, 5);
    printf(", \"dmma_dep_chain_ns\": %.2f", ms * 1e6 / iters);
    ms = time_ms([&] { dfma_kernel<1><<<blocks, threads>>>(out, iters, 0.999, 1e-3); }, 5);
    printf(", \"dfma_dep_chain_ns\": %.2f", ms * 1e6 / iters);
    ms = time_ms([&] { dmma_kernel<4><<<blocks, threads>>>(out, iters); }, 5);
    printf(", \"dmma_4chain_1wpsmsp_tflops\": %.2f", 2.0 * 256 * 4 * iters * (double)blocks * threads / 32 / ms * 1e-9);
    ms = time_ms([&] { dmma_kernel<16><<<blocks, threads>>>(out, iters); }, 5);
    printf(", \"dmma_16chain_1wpsmsp_tflops\": %.2f", 2.0 * 256 * 16 * iters * (double)blocks * threads / 32 / ms * 1e-9);
  }
  {
    int blocks = sms * 8, threads = 256, iters = 2000;
    double ms = time_ms([&] { exp_kernel<<<blocks, threads>>>(out, iters, 1e-4); }, 5);
    printf(", \"exp_gops\": %.1f", 4.0 * iters * (double)blocks * threads / ms * 1e-6);
  }
  cublasHandle_t h; cublasCreate(&h);
  for (int n : {2048, 4096, 8192}) {
    double *A, *B, *C; size_t bytes = sizeof(double) * n * n;
    CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
    CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes)); CK(cudaMemset(C, 0, bytes));
    double al = 1.0, be = 0.0;
    double ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &al, A, n, B, n, &be, C, n); }, 10);
    printf(", \"dgemm_nt_%d_tflops\": %.2f", n, 2.0 * n * n * (double)n / ms * 1e-9);
    ms = time_ms([&] { cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, n, &al, A, n, &be, C, n); }, 10);
    printf(", \"dsyrk_%d_tflops\": %.2f", n, 1.0 * n * n * (double)n / ms * 1e-9);
    cudaFree(A); cudaFree(B); cudaFree(C);
  }
  {  // sustained DGEMM (≈3 s) to see the power-capped figure
    int n = 8192; double *A, *B, *C; size_t bytes = sizeof(double) * n * n;
    CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
    CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes));
    double al = 1.0, be = 0.0;
    double ms = time_ms([&] { for (int r = 0; r < 60; r++) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &al, A, n, B, n, &be, C, n); }, 2);
    printf(", \"dgemm_nt_8192_sustained_tflops\": %.2f", 60 * 2.0 * n * n * (double)n / ms * 1e-9);
    cudaFree(A); cudaFree(B); cudaFree(C);
  }
  {  // HBM copy
    size_t n = (size_t)1 << 28; double *a, *b; CK(cudaMalloc(&a, n * 8)); CK(cudaMalloc(&b, n * 8));
    double ms = time_ms([&] { cudaMemcpyAsync(b, a, n * 8, cudaMemcpyDeviceToDevice); }, 10);
    printf(", \"hbm_copy_gbs\": %.1f", 2.0 * n * 8 / ms * 1e-6);
    cudaFree(a); cudaFree(b);
  }
  printf("}\n");
  return 0;
}
