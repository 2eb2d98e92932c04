// Is FP64 DFMA issue limited by register-file operand bandwidth when all three 64-bit operands are
// distinct registers (the matrix-in-registers mat-vec)?  nvcc -arch=sm_100a -O3 -o ubench_rf ubench_rf.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int ROWS>
__global__ void __launch_bounds__(256, 1) k(double *out, long long *cyc, int iters, const double *in) {
  double S[ROWS][42 / ROWS];
  double r[6];
#pragma unroll
  for (int q = 0; q < ROWS; ++q)
#pragma unroll
    for (int j = 0; j < 42 / ROWS; ++j) S[q][j] = in[(threadIdx.x + j + 50 * q) & 255];
#pragma unroll
  for (int j = 0; j < 6; ++j) r[j] = in[(threadIdx.x * 3 + j) & 255];
  double acc[4] = {0, 0, 0, 0};
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {            // 3 distinct operands: S[j], r[j%6], acc[j%4]
#pragma unroll
      for (int j = 0; j < 42; ++j) acc[j & 3] = fma(S[0][j % (42 / ROWS)], r[j % 6], acc[j & 3]);
    } else if (MODE == 1) {     // second operand fixed -> reuse cache
#pragma unroll
      for (int j = 0; j < 42; ++j) acc[j & 3] = fma(S[0][j % (42 / ROWS)], r[0], acc[j & 3]);
    } else {                    // ROWS rows share each r value: consecutive DFMAs reuse r
#pragma unroll
      for (int j = 0; j < 42 / ROWS; ++j)
#pragma unroll
        for (int q = 0; q < ROWS; ++q) acc[q & 3] = fma(S[q][j], r[j % 6], acc[q & 3]);
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) r[j] += 1e-9;     // keep the loop body from being hoisted
  }
  const long long t1 = clock64();
  if (acc[0] + acc[1] + acc[2] + acc[3] == 123.456) out[threadIdx.x] = acc[0];
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  double *out, *in; long long *cyc, h;
  cudaMalloc(&out, 4096); cudaMalloc(&in, 4096); cudaMalloc(&cyc, 8);
  double hin[256];
  for (int i = 0; i < 256; ++i) hin[i] = 1e-3 * i;
  cudaMemcpy(in, hin, 2048, cudaMemcpyHostToDevice);
  const int iters = 4000;
#define RUN(name, threads, ...)                                                       \
  k<__VA_ARGS__><<<1, threads>>>(out, cyc, iters, in); cudaDeviceSynchronize();       \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                     \
  printf("%-58s %d thr: %7.1f cycles per 42 DFMA (+6 DADD)\n", name, threads, (double)h / iters);
  RUN("S[j]*r[j%6]+acc  (3 distinct regs)", 32, 0, 1);
  RUN("S[j]*r[j%6]+acc  (3 distinct regs)", 128, 0, 1);
  RUN("S[j]*r[j%6]+acc  (3 distinct regs)", 256, 0, 1);
  RUN("S[j]*r0+acc      (one operand fixed)", 32, 1, 1);
  RUN("S[j]*r0+acc      (one operand fixed)", 256, 1, 1);
  RUN("2 rows share r_j (pairs of DFMA reuse r)", 32, 2, 2);
  RUN("2 rows share r_j (pairs of DFMA reuse r)", 256, 2, 2);
  RUN("3 rows share r_j", 256, 2, 3);
  return 0;
}
