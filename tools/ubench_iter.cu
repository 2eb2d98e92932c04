// Synthetic copy of the team kernel's ADMM iteration (3 barrier-separated phases) to find what its
// latency floor is on B200.  nvcc -arch=sm_100a -O3 -o ubench_iter ubench_iter.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 81, M = 136, HALF = 42, NP2 = 84, MP = 136, NCT = 192;

template <int WA, int WAT2, bool DO_AT, bool DO_S, bool DO_A, int NBAR>
__global__ void __launch_bounds__(256, 1) iter_kernel(double *out, long long *cyc, int iters, const int *perm) {
  __shared__ __align__(16) double vbuf[MP], rbuf[NP2], xtbuf[NP2], lob[MP], hib[MP], rinvb[MP];
  const int tid = threadIdx.x, half = tid & 1, pairi = tid >> 1;
  const bool has_col = pairi < N, col_warp = tid < NCT, has_row = tid < M;
  double S[HALF], Ar[8], ATr[8];
  int Aoff[8], AToff[8];
#pragma unroll
  for (int j = 0; j < HALF; ++j) S[j] = 1e-3 * (j + 1) + 1e-6 * tid;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    Ar[e] = 0.01 * (e + 1); ATr[e] = 0.02 * (e + 1);
    Aoff[e] = 8 * (perm[(tid * 8 + e) % 1024] % N);
    AToff[e] = 8 * (perm[(tid * 8 + e + 512) % 1024] % M);
  }
  if (tid < MP) { vbuf[tid] = 0.1; lob[tid] = -1; hib[tid] = 1; rinvb[tid] = 10.0; }
  if (tid < NP2) { rbuf[tid] = 0.0; xtbuf[tid] = 0.0; }
  double x = 0.1, z = 0.0, y = 0.0;
  const double rv = 0.1, alpha = 1.6, oma = -0.6, sigma = 1e-6;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (DO_AT && col_warp) {
      double g[WAT2], a0 = 0, a1 = 0;
#pragma unroll
      for (int e = 0; e < WAT2; ++e) g[e] = *reinterpret_cast<const double *>(reinterpret_cast<const unsigned char *>(vbuf) + AToff[e]);
#pragma unroll
      for (int e = 0; e < WAT2; ++e) { if (e & 1) a1 = fma(ATr[e], g[e], a1); else a0 = fma(ATr[e], g[e], a0); }
      double s = a0 + a1;
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (has_col && half == 0) rbuf[pairi] = sigma * x - 0.5 + s;
    }
    if (NBAR >= 1) __syncthreads();
    if (DO_S && col_warp) {
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      const double2 *r2 = reinterpret_cast<const double2 *>(rbuf + half * HALF);
#pragma unroll
      for (int j = 0; j < HALF / 2; ++j) {
        const double2 rr = r2[j];
        if (j & 1) { a2 = fma(S[2 * j], rr.x, a2); a3 = fma(S[2 * j + 1], rr.y, a3); }
        else { a0 = fma(S[2 * j], rr.x, a0); a1 = fma(S[2 * j + 1], rr.y, a1); }
      }
      double xt = (a0 + a1) + (a2 + a3);
      xt += __shfl_xor_sync(0xffffffffu, xt, 1);
      if (has_col && half == 0) xtbuf[pairi] = xt * 1e-3;
      x = alpha * xt * 1e-3 + oma * x;
    }
    if (NBAR >= 2) __syncthreads();
    if (DO_A && has_row) {
      double g[WA], a0 = 0, a1 = 0;
#pragma unroll
      for (int e = 0; e < WA; ++e) g[e] = *reinterpret_cast<const double *>(reinterpret_cast<const unsigned char *>(xtbuf) + Aoff[e]);
#pragma unroll
      for (int e = 0; e < WA; ++e) { if (e & 1) a1 = fma(Ar[e], g[e], a1); else a0 = fma(Ar[e], g[e], a0); }
      const double zt = a0 + a1;
      const double zr = alpha * zt + oma * z;
      const double zn = fmin(fmax(zr + rinvb[tid] * y, lob[tid]), hib[tid]);
      const double dy = rv * (zr - zn);
      y += dy;
      z = zn;
      vbuf[tid] = rv * zn - y;
    }
    if (NBAR >= 3) __syncthreads();
  }
  const long long t1 = clock64();
  if (x + z + y == 123.456) out[tid] = x;
  if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  double *out; long long *cyc, h; int *perm;
  cudaMalloc(&out, 1 << 16); cudaMalloc(&cyc, 8); cudaMalloc(&perm, 4096);
  int hp[1024];
  unsigned s = 12345;
  for (int i = 0; i < 1024; ++i) { s = s * 1664525u + 1013904223u; hp[i] = (s >> 8) % 1000; }
  cudaMemcpy(perm, hp, 4096, cudaMemcpyHostToDevice);
  const int iters = 4000;
#define RUN(name, ...)                                                                        \
  iter_kernel<__VA_ARGS__><<<1, 256>>>(out, cyc, iters, perm); cudaDeviceSynchronize();       \
  iter_kernel<__VA_ARGS__><<<1, 256>>>(out, cyc, iters, perm); cudaDeviceSynchronize();       \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                             \
  printf("%-52s %8.1f cycles / iteration   (%s)\n", name, (double)h / iters, cudaGetErrorString(cudaGetLastError()));
  RUN("full iteration (WA=8, WAT2=6), 3 barriers", 8, 6, true, true, true, 3);
  RUN("only barriers", 8, 6, false, false, false, 3);
  RUN("AT phase + barriers", 8, 6, true, false, false, 3);
  RUN("S phase + barriers", 8, 6, false, true, false, 3);
  RUN("A phase + barriers", 8, 6, false, false, true, 3);
  RUN("S phase alone, 1 barrier", 8, 6, false, true, false, 1);
  RUN("full, WA=4, WAT2=3", 4, 3, true, true, true, 3);
  RUN("full, WA=1, WAT2=1", 1, 1, true, true, true, 3);
  // several teams per SM: does the per-iteration time hold when 148 blocks run (one per SM)?
  iter_kernel<8, 6, true, true, true, 3><<<148, 256>>>(out, cyc, iters, perm); cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-52s %8.1f cycles / iteration\n", "full iteration, grid 148", (double)h / iters);
  return 0;
}
