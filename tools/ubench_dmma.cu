// Micro-benchmark that sizes the multi-RHS DMMA path (csrc/wave.cuh): the two spectral GEMMs
//     T = R V' (8 or 16 lanes x n) -> W = T .* d -> X = W V'^T
// chained through registers (the C fragment of GEMM 1 is the A fragment of GEMM 2 once the spectral index is
// ordered J(t, c, e) = 8t + 4e + c), V' read from shared memory as B fragments (one table, stride = 4 mod 16).
// Reports cycles per warp-iteration and DMMA issue rate per SM for 4 / 8 / 12 / 16 warps per SM, and checks the
// result of one iteration against a scalar loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_dmma ubench_dmma.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int N = 81;
constexpr int KS = (N + 3) / 4;       // 21 k-steps over positions (GEMM 1)
constexpr int NT = (N + 7) / 8;       // 11 tiles of 8 (spectral tiles in GEMM 1, position tiles in GEMM 2)
constexpr int LD = 84;                // 84 = 4 (mod 16)
constexpr int ROWS = 8 * NT;          // 88

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MT = lanes tiles per warp (1: 8 lanes, 2: 16 lanes sharing every B fragment)
template <int MT, int MAXW>
__global__ void __launch_bounds__(32 * MAXW, 1) gemm_chain(const double *Vg, const double *Rg, const double *dg, double *Xg, long long *cyc,
                                                     int iters, int do_check) {
  extern __shared__ __align__(16) double sm[];
  double *V = sm;                                   // [ROWS][LD]: V[p][J]
  const int tid = threadIdx.x, warp = tid >> 5, lid = tid & 31, g = lid >> 2, c = lid & 3;
  for (int i = tid; i < ROWS * LD; i += blockDim.x) V[i] = Vg[i];
  double *nb = sm + ROWS * LD + (size_t)warp * MT * 8 * LD;      // per warp: MT*8 lanes x LD
  for (int i = lid; i < MT * 8 * LD; i += 32) nb[i] = Rg[(size_t)(warp * MT * 8) * LD + i];
  double dsc[MT][2 * NT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int e = 0; e < 2; ++e) dsc[mt][2 * t + e] = dg[(size_t)(warp * MT * 8 + mt * 8 + g) * ROWS + 8 * t + 4 * e + c];
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // A fragments of GEMM 1: r[lane g][p = 4s + c]
    double af[MT][KS];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int s = 0; s < KS; ++s) af[mt][s] = nb[(mt * 8 + g) * LD + 4 * s + c];
    __syncwarp();
    double acc[MT][2 * NT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int j = 0; j < 2 * NT; ++j) acc[mt][j] = 0.0;
    // GEMM 1: T[lane][J] = sum_p r[lane][p] V[p][J];  B fragment: thread (krow = c, ncol = g) <-> V[4s + c][8t + 4(g&1) + (g>>1)]
#pragma unroll
    for (int s = 0; s < KS; ++s) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const double b = V[(4 * s + c) * LD + 8 * t + 4 * (g & 1) + (g >> 1)];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) dmma(acc[mt][2 * t], acc[mt][2 * t + 1], af[mt][s], b);
      }
    }
    // scale: thread (g, c) holds T[lane g][J = 8t + 4e + c] in acc[2t + e]
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int j = 0; j < 2 * NT; ++j) acc[mt][j] *= dsc[mt][j];
    // GEMM 2: X[lane][p] = sum_J W[lane][J] V[p][J]; k-step (2t + e) covers J = 8t + 4e + {0..3}: its A fragment is acc[2t + e];
    // B fragment: thread (krow = c, ncol = g) <-> V[8t' + g][8t + 4e + c]
#pragma unroll
    for (int tp = 0; tp < NT; ++tp) {
      double x0[MT], x1[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) x0[mt] = x1[mt] = 0.0;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {           // J = 4ks + c < 84: the last half tile of GEMM 1 is padding
        const double b = V[(8 * tp + g) * LD + 4 * ks + c];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) dmma(x0[mt], x1[mt], acc[mt][ks], b);
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int p = 8 * tp + 2 * c;
        if (p < LD) nb[(mt * 8 + g) * LD + p] = x0[mt] * 1e-3;       // feed back (scaled down so the loop stays finite)
        if (p + 1 < LD) nb[(mt * 8 + g) * LD + p + 1] = x1[mt] * 1e-3;
      }
    }
    __syncwarp();
  }
  const long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  if (do_check && blockIdx.x == 0)
    for (int i = lid; i < MT * 8 * LD; i += 32) Xg[(size_t)(warp * MT * 8) * LD + i] = nb[i];
}

int main() {
  const int maxl = 16 * 16;
  std::vector<double> V(ROWS * LD, 0.0), R((size_t)maxl * LD, 0.0), d((size_t)maxl * ROWS, 0.0), X((size_t)maxl * LD);
  srand(1);
  for (int p = 0; p < N; ++p)
    for (int j = 0; j < N; ++j) V[p * LD + j] = (rand() / (double)RAND_MAX - 0.5) * 0.2;
  for (int l = 0; l < maxl; ++l) {
    for (int p = 0; p < N; ++p) R[(size_t)l * LD + p] = rand() / (double)RAND_MAX - 0.5;
    for (int j = 0; j < N; ++j) d[(size_t)l * ROWS + j] = 1.0 / (1.0 + 0.1 * (l + 1) * j);
  }
  double *dV, *dR, *dd, *dX;
  long long *dc, hc;
  cudaMalloc(&dV, V.size() * 8); cudaMalloc(&dR, R.size() * 8); cudaMalloc(&dd, d.size() * 8); cudaMalloc(&dX, X.size() * 8);
  cudaMalloc(&dc, 8);
  cudaMemcpy(dV, V.data(), V.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dR, R.data(), R.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dd, d.data(), d.size() * 8, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(gemm_chain<1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(gemm_chain<1, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(gemm_chain<2, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(gemm_chain<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  // ---- correctness of one iteration (MT = 1 and 2)
  for (int mt = 1; mt <= 2; ++mt) {
    const int warps = 4, lanes = warps * 8 * mt;
    const size_t smem = (size_t)(ROWS * LD + warps * mt * 8 * LD) * 8;
    if (mt == 1) gemm_chain<1, 16><<<1, 32 * warps, smem>>>(dV, dR, dd, dX, dc, 1, 1);
    else gemm_chain<2, 8><<<1, 32 * warps, smem>>>(dV, dR, dd, dX, dc, 1, 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(X.data(), dX, X.size() * 8, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int l = 0; l < lanes; ++l) {
      double T[ROWS] = {0};
      for (int j = 0; j < N; ++j) {
        double s = 0;
        for (int p = 0; p < N; ++p) s += R[(size_t)l * LD + p] * V[p * LD + j];
        T[j] = s * d[(size_t)l * ROWS + j];
      }
      for (int p = 0; p < N; ++p) {
        double s = 0;
        for (int j = 0; j < N; ++j) s += T[j] * V[p * LD + j];
        worst = fmax(worst, fabs(s * 1e-3 - X[(size_t)l * LD + p]));
      }
    }
    printf("check MT=%d: max |x - ref| = %.3e over %d lanes %s\n", mt, worst, lanes, worst < 1e-15 ? "OK" : "MISMATCH");
  }
  // ---- timing: one CTA per SM, W warps
  const int iters = 200, dm1 = 2 * KS * NT;       // DMMA per warp-iteration per lanes tile
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  for (int mt = 1; mt <= 2; ++mt)
    for (int warps : {1, 4, 8, 12, 16}) {
      if (mt == 2 && warps > 12) continue;
      const size_t smem = (size_t)(ROWS * LD + warps * mt * 8 * LD) * 8;
      if (smem > 220 * 1024) continue;
      for (int rep = 0; rep < 2; ++rep) {
        if (mt == 1 && warps > 12) gemm_chain<1, 16><<<nsm, 32 * warps, smem>>>(dV, dR, dd, dX, dc, iters, 0);
        else if (mt == 1) gemm_chain<1, 12><<<nsm, 32 * warps, smem>>>(dV, dR, dd, dX, dc, iters, 0);
        else if (warps > 8) gemm_chain<2, 12><<<nsm, 32 * warps, smem>>>(dV, dR, dd, dX, dc, iters, 0);
        else gemm_chain<2, 8><<<nsm, 32 * warps, smem>>>(dV, dR, dd, dX, dc, iters, 0);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
      const double cpi = (double)hc / iters;
      const double dmma_per_cyc = (double)dm1 * mt * warps / cpi;
      printf("MT=%d (%2d lanes/warp) warps/SM=%2d: %8.0f cycles per warp-iteration, %.3f DMMA/cycle/SM (pipe peak 0.25) -> %.1f %% ; %.2f lane-iterations/kcycle/SM\n",
             mt, 8 * mt, warps, cpi, dmma_per_cyc, 100.0 * dmma_per_cyc / 0.25, 1000.0 * warps * mt * 8 / cpi);
    }
  printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
