// FP64 roofline denominators measured on the device the bench runs on.
// MEASURED_PEAKS.json carries only HBM GB/s and bf16 TFLOP/s; this path computes in float64
// (DFMA on the FP64 pipe, DMMA mma.sync m8n8k4 on the tensor pipe; tcgen05 has no f64 kind),
// so bench.py measures both rates live and reports against them.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mpcb.h"

__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double seed) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x * 1e-3;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(double *out, int iters, double seed) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = seed + threadIdx.x * 1e-3, b = 1e-3 * (threadIdx.x & 7);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int mpcb_measure_fp64_peak(int device, int use_dmma, double *tflops) {
  if (!tflops) return MPCB_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return MPCB_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MPCB_ERR_CUDA;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
  double *d = nullptr;
  if (cudaMalloc(&d, (size_t)blocks * threads * 8) != cudaSuccess) return MPCB_ERR_NOMEM;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    if (use_dmma) dmma_peak_kernel<<<blocks, threads>>>(d, iters, 1.0 + rep);
    else dfma_peak_kernel<<<blocks, threads>>>(d, iters, 1.0 + rep);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    // flops: DFMA = 2 per thread-op; DMMA m8n8k4 = 2*8*8*4 = 512 per warp-op
    const double ops = use_dmma ? (double)blocks * (threads / 32) * iters * 8.0 * 512.0
                                : (double)blocks * threads * iters * 8.0 * 2.0;
    if (rep > 0) best = ops / (ms * 1e-3) / 1e12 > best ? ops / (ms * 1e-3) / 1e12 : best;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  if (cudaGetLastError() != cudaSuccess) return MPCB_ERR_CUDA;
  *tflops = best;
  return MPCB_OK;
}
