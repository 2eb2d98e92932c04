// Micro-benchmarks that size the team kernel: FP64 DFMA latency / issue rate per warp on B200,
// broadcast LDS.128 rate, block barrier cost.  Build: nvcc -arch=sm_100a -O3 -o ubench_fp64 ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_chain(double *out, long long *cyc, int iters) {
  double a[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) a[i] = 1.0 + threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], b, c);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += a[i];
  if (s == 123.456) out[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void lds_bcast(double *out, long long *cyc, int iters) {
  __shared__ __align__(16) double buf[96];
  if (threadIdx.x < 96) buf[threadIdx.x] = threadIdx.x;
  __syncthreads();
  const double2 *b2 = reinterpret_cast<const double2 *>(buf);
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  const double m = 1.0 + threadIdx.x * 1e-9;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 20; j += 2) {
      const double2 r0 = b2[j], r1 = b2[j + 1];
      a0 = fma(m, r0.x, a0); a1 = fma(m, r0.y, a1); a2 = fma(m, r1.x, a2); a3 = fma(m, r1.y, a3);
    }
  }
  const long long t1 = clock64();
  if (a0 + a1 + a2 + a3 == 123.456) out[threadIdx.x] = a0;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void bar_cost(long long *cyc, int iters) {
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 4096;
#define RUN(name, kern, blocks, threads, per)                                    \
  kern<<<blocks, threads>>>(out, cyc, iters); cudaDeviceSynchronize();           \
  kern<<<blocks, threads>>>(out, cyc, iters); cudaDeviceSynchronize();           \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                \
  printf("%-44s %8.2f cycles per %s\n", name, (double)h / iters, per);
  RUN("DFMA 1 chain, 1 warp", dfma_chain<1>, 1, 32, "dependent DFMA");
  RUN("DFMA 2 chains, 1 warp", dfma_chain<2>, 1, 32, "2 DFMA");
  RUN("DFMA 4 chains, 1 warp", dfma_chain<4>, 1, 32, "4 DFMA");
  RUN("DFMA 8 chains, 1 warp", dfma_chain<8>, 1, 32, "8 DFMA");
  RUN("DFMA 4 chains, 4 warps (1/SMSP)", dfma_chain<4>, 1, 128, "4 DFMA");
  RUN("DFMA 4 chains, 8 warps (2/SMSP)", dfma_chain<4>, 1, 256, "4 DFMA");
  RUN("DFMA 4 chains, 16 warps (4/SMSP)", dfma_chain<4>, 1, 512, "4 DFMA");
  RUN("10x(LDS.128 bcast + 4 DFMA... ) 1 warp", lds_bcast, 1, 32, "20 LDS.128 + 40 DFMA");
  RUN("same, 8 warps", lds_bcast, 1, 256, "20 LDS.128 + 40 DFMA");
  RUN("same, 16 warps", lds_bcast, 1, 512, "20 LDS.128 + 40 DFMA");
#define RUNB(threads)                                                            \
  bar_cost<<<1, threads>>>(cyc, iters); cudaDeviceSynchronize();                 \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                \
  printf("__syncthreads, %d threads: %.2f cycles\n", threads, (double)h / iters);
  RUNB(32) RUNB(96) RUNB(128) RUNB(256) RUNB(512)
  return 0;
}
