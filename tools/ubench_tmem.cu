// Can tensor memory hold the per-lane operator?  Synthetic team-kernel iteration (as ubench_iter.cu) with the
// 42 doubles of S each thread owns kept (a) in registers (2 teams per SM is the register-file limit) or
// (b) 32 of them in TMEM (64 32-bit columns per thread, tcgen05.ld.32x32b.x32 twice per iteration) and 10
// in shared memory, which frees the registers for 3 or 4 teams per SM.  Prints aggregate lane-iterations
// per microsecond per SM for each arrangement.   nvcc -arch=sm_100a -O3 -o ubench_tmem ubench_tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int N = 81, M = 136, HALF = 42, NP2 = 84, MP = 136, NCT = 192;

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
#define TMEM_WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}

template <int TM, int CTAS>
__global__ void __launch_bounds__(256, CTAS) iter_kernel(double *out, long long *cyc, int iters, const int *perm, int structured) {
  __shared__ __align__(16) double vbuf[MP], rbuf[NP2], xtbuf[NP2], lob[MP], hib[MP], rinvb[MP];
  constexpr bool TMEM = TM > 0;
  constexpr int SSM = TMEM ? HALF - TM : 0, TCOLS = (TM == 42) ? 256 : 128, TSTRIDE = (TM == 42) ? 84 : 64;
  __shared__ __align__(16) double Ssm[SSM > 0 ? SSM * 256 : 1];
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, half = tid & 1, pairi = tid >> 1, warp = tid >> 5;
  const bool has_col = pairi < N, col_warp = tid < NCT, has_row = tid < M;
  double S[TMEM ? 1 : HALF], Ar[8], ATr[8];
  int Aoff[8], AToff[8];
  uint32_t taddr = 0;
  if (TMEM) {
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase_s)), "n"(TCOLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    taddr = tbase_s + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(TSTRIDE * (warp >> 2));
    for (int hseg = 0; hseg < (TM == 42 ? 3 : 2); ++hseg) {
      uint32_t r[32];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const double v = 1e-3 * (16 * hseg + j + 1) + 1e-6 * tid;
        r[2 * j] = (uint32_t)__double2loint(v);
        r[2 * j + 1] = (uint32_t)__double2hiint(v);
      }
      if (hseg < 2) tmem_st32(taddr + 32 * hseg, r);
      else { uint32_t r16[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) r16[j] = r[j];
        tmem_st16(taddr + 64, r16); uint32_t r4[4] = {r[16], r[17], r[18], r[19]}; tmem_st4(taddr + 80, r4); }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    for (int j = 0; j < SSM; ++j) Ssm[j * 256 + tid] = 1e-3 * (TM + j + 1) + 1e-6 * tid;
  } else {
#pragma unroll
    for (int j = 0; j < HALF; ++j) S[j] = 1e-3 * (j + 1) + 1e-6 * tid;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    Ar[e] = 0.01 * (e + 1); ATr[e] = 0.02 * (e + 1);
    Aoff[e] = structured ? 8 * ((tid + 5 * e) % N) : 8 * (perm[(tid * 8 + e) % 1024] % N);
    AToff[e] = structured ? 8 * ((tid / 2 + 7 * e + 3 * (tid & 1)) % M) : 8 * (perm[(tid * 8 + e + 512) % 1024] % M);
  }
  if (tid < MP) { vbuf[tid] = 0.1; lob[tid] = -1; hib[tid] = 1; rinvb[tid] = 10.0; }
  if (tid < NP2) { rbuf[tid] = 0.0; xtbuf[tid] = 0.0; }
  double x = 0.1, z = 0.0, y = 0.0;
  const double rv = 0.1, alpha = 1.6, oma = -0.6, sigma = 1e-6;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (col_warp) {
      double g[6], a0 = 0, a1 = 0;
#pragma unroll
      for (int e = 0; e < 6; ++e) g[e] = *reinterpret_cast<const double *>(reinterpret_cast<const unsigned char *>(vbuf) + AToff[e]);
#pragma unroll
      for (int e = 0; e < 6; ++e) { if (e & 1) a1 = fma(ATr[e], g[e], a1); else a0 = fma(ATr[e], g[e], a0); }
      double s = a0 + a1;
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (has_col && half == 0) rbuf[pairi] = sigma * x - 0.5 + s;
    }
    __syncthreads();
    if (col_warp) {
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      const double2 *r2 = reinterpret_cast<const double2 *>(rbuf + half * HALF);
      if (TMEM) {
        uint32_t ca[16], cb[16];
        auto use = [&](const uint32_t (&c)[16], int q) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const double2 rr = r2[4 * q + j];
            const double sa = __hiloint2double((int)c[4 * j + 1], (int)c[4 * j]), sb = __hiloint2double((int)c[4 * j + 3], (int)c[4 * j + 2]);
            if (j & 1) { a2 = fma(sa, rr.x, a2); a3 = fma(sb, rr.y, a3); } else { a0 = fma(sa, rr.x, a0); a1 = fma(sb, rr.y, a1); }
          }
        };
        tmem_ld16(taddr, ca);
        tmem_ld16(taddr + 16, cb);
        TMEM_WAIT_LD();
        use(ca, 0);
        tmem_ld16(taddr + 32, ca);
        use(cb, 1);
        tmem_ld16(taddr + 48, cb);
        TMEM_WAIT_LD();
        use(ca, 2);
        if (TM == 42) tmem_ld16(taddr + 64, ca);
        use(cb, 3);
        if (TM == 42) {
          uint32_t c4[4];
          tmem_ld4(taddr + 80, c4);
          TMEM_WAIT_LD();
          use(ca, 4);
          const double2 rr = r2[20];
          a0 = fma(__hiloint2double((int)c4[1], (int)c4[0]), rr.x, a0);
          a1 = fma(__hiloint2double((int)c4[3], (int)c4[2]), rr.y, a1);
        }
#pragma unroll
        for (int j = 0; j < SSM / 2; ++j) {
          const double2 rr = r2[16 + j];
          a0 = fma(Ssm[(2 * j) * 256 + tid], rr.x, a0);
          a1 = fma(Ssm[(2 * j + 1) * 256 + tid], rr.y, a1);
        }
      } else {
#pragma unroll
        for (int j = 0; j < HALF / 2; ++j) {
          const double2 rr = r2[j];
          if (j & 1) { a2 = fma(S[2 * j], rr.x, a2); a3 = fma(S[2 * j + 1], rr.y, a3); }
          else { a0 = fma(S[2 * j], rr.x, a0); a1 = fma(S[2 * j + 1], rr.y, a1); }
        }
      }
      double xt = (a0 + a1) + (a2 + a3);
      xt += __shfl_xor_sync(0xffffffffu, xt, 1);
      if (has_col && half == 0) xtbuf[pairi] = xt * 1e-3;
      x = alpha * xt * 1e-3 + oma * x;
    }
    __syncthreads();
    if (has_row) {
      double g[8], a0 = 0, a1 = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] = *reinterpret_cast<const double *>(reinterpret_cast<const unsigned char *>(xtbuf) + Aoff[e]);
#pragma unroll
      for (int e = 0; e < 8; ++e) { if (e & 1) a1 = fma(Ar[e], g[e], a1); else a0 = fma(Ar[e], g[e], a0); }
      const double zt = a0 + a1;
      const double zr = alpha * zt + oma * z;
      const double zn = fmin(fmax(zr + rinvb[tid] * y, lob[tid]), hib[tid]);
      const double dy = rv * (zr - zn);
      y += dy;
      z = zn;
      vbuf[tid] = rv * zn - y;
    }
    __syncthreads();
  }
  const long long t1 = clock64();
  if (x + z + y == 123.456) out[tid] = x;
  if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  if (TMEM) {
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase_s), "n"(TCOLS));
  }
}


// ---- NL trajectories interleaved phase by phase inside ONE 256-thread team (1 team per SM): barriers and dependent chains
//      are shared NL ways; S of every slot in TMEM (NL*84 columns per thread), sparse rows in registers (shared by the slots).
template <int NL>
__global__ void __launch_bounds__(256, 1) slots_kernel(double *out, long long *cyc, int iters, const int *perm, int structured) {
  __shared__ __align__(16) double vbuf[NL][MP], rbuf[NL][NP2], xtbuf[NL][NP2], lob[MP], hib[MP], rinvb[MP];
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, half = tid & 1, pairi = tid >> 1, warp = tid >> 5;
  const bool has_col = pairi < N, col_warp = tid < NCT, has_row = tid < M;
  double Ar[8], ATr[8];
  int Aoff[8], AToff[8];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = tbase_s + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(NL * 84 * (warp >> 2));
  for (int l = 0; l < NL; ++l)
    for (int g = 0; g < 5; ++g) {
      uint32_t r16[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double v = 1e-3 * (8 * g + j + 1) + 1e-6 * tid + 1e-5 * l;
        r16[2 * j] = (uint32_t)__double2loint(v);
        r16[2 * j + 1] = (uint32_t)__double2hiint(v);
      }
      tmem_st16(taddr + 84 * l + 16 * g, r16);
    }
  for (int l = 0; l < NL; ++l) { uint32_t r4[4] = {0, 0x3f500000u, 0, 0x3f500000u}; tmem_st4(taddr + 84 * l + 80, r4); }
  asm volatile("tcgen05.wait::st.sync.aligned;");
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    Ar[e] = 0.01 * (e + 1); ATr[e] = 0.02 * (e + 1);
    Aoff[e] = structured ? 8 * ((tid + 5 * e) % N) : 8 * (perm[(tid * 8 + e) % 1024] % N);
    AToff[e] = structured ? 8 * ((tid / 2 + 7 * e + 3 * (tid & 1)) % M) : 8 * (perm[(tid * 8 + e + 512) % 1024] % M);
  }
  if (tid < MP) { for (int l = 0; l < NL; ++l) vbuf[l][tid] = 0.1; lob[tid] = -1; hib[tid] = 1; rinvb[tid] = 10.0; }
  if (tid < NP2) for (int l = 0; l < NL; ++l) { rbuf[l][tid] = 0.0; xtbuf[l][tid] = 0.0; }
  double x[NL], z[NL], y[NL];
#pragma unroll
  for (int l = 0; l < NL; ++l) { x[l] = 0.1 + 0.01 * l; z[l] = 0.0; y[l] = 0.0; }
  const double rv = 0.1, alpha = 1.6, oma = -0.6, sigma = 1e-6;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (col_warp) {
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        double g[6], a0 = 0, a1 = 0;
#pragma unroll
        for (int e = 0; e < 6; ++e) g[e] = *reinterpret_cast<const double *>(reinterpret_cast<const unsigned char *>(vbuf[l]) + AToff[e]);
#pragma unroll
        for (int e = 0; e < 6; ++e) { if (e & 1) a1 = fma(ATr[e], g[e], a1); else a0 = fma(ATr[e], g[e], a0); }
        double s = a0 + a1;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (has_col && half == 0) rbuf[l][pairi] = sigma * x[l] - 0.5 + s;
      }
    }
    __syncthreads();
    if (col_warp) {
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        const double2 *r2 = reinterpret_cast<const double2 *>(rbuf[l] + half * HALF);
        uint32_t ca[16], cb[16];
        auto use = [&](const uint32_t (&c)[16], int q) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const double2 rr = r2[4 * q + j];
            const double sa = __hiloint2double((int)c[4 * j + 1], (int)c[4 * j]), sb = __hiloint2double((int)c[4 * j + 3], (int)c[4 * j + 2]);
            if (j & 1) { a2 = fma(sa, rr.x, a2); a3 = fma(sb, rr.y, a3); } else { a0 = fma(sa, rr.x, a0); a1 = fma(sb, rr.y, a1); }
          }
        };
        const uint32_t ta = taddr + 84 * l;
        tmem_ld16(ta, ca); tmem_ld16(ta + 16, cb); TMEM_WAIT_LD();
        use(ca, 0); tmem_ld16(ta + 32, ca);
        use(cb, 1); tmem_ld16(ta + 48, cb); TMEM_WAIT_LD();
        use(ca, 2); tmem_ld16(ta + 64, ca);
        use(cb, 3);
        uint32_t c4[4];
        tmem_ld4(ta + 80, c4); TMEM_WAIT_LD();
        use(ca, 4);
        const double2 rr = r2[20];
        a0 = fma(__hiloint2double((int)c4[1], (int)c4[0]), rr.x, a0);
        a1 = fma(__hiloint2double((int)c4[3], (int)c4[2]), rr.y, a1);
        double xt = (a0 + a1) + (a2 + a3);
        xt += __shfl_xor_sync(0xffffffffu, xt, 1);
        if (has_col && half == 0) xtbuf[l][pairi] = xt * 1e-3;
        x[l] = alpha * xt * 1e-3 + oma * x[l];
      }
    }
    __syncthreads();
    if (has_row) {
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        double g[8], a0 = 0, a1 = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] = *reinterpret_cast<const double *>(reinterpret_cast<const unsigned char *>(xtbuf[l]) + Aoff[e]);
#pragma unroll
        for (int e = 0; e < 8; ++e) { if (e & 1) a1 = fma(Ar[e], g[e], a1); else a0 = fma(Ar[e], g[e], a0); }
        const double zt = a0 + a1;
        const double zr = alpha * zt + oma * z[l];
        const double zn = fmin(fmax(zr + rinvb[tid] * y[l], lob[tid]), hib[tid]);
        const double dy = rv * (zr - zn);
        y[l] += dy;
        z[l] = zn;
        vbuf[l][tid] = rv * zn - y[l];
      }
    }
    __syncthreads();
  }
  const long long t1 = clock64();
  double acc = 0;
#pragma unroll
  for (int l = 0; l < NL; ++l) acc += x[l] + z[l] + y[l];
  if (acc == 123.456) out[tid] = acc;
  if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase_s));
}

template <int NL>
static void run_slots(double *out, long long *cyc, const int *perm, int structured) {
  const int iters = 4000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  slots_kernel<NL><<<148, 256>>>(out, cyc, iters, perm, structured);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  slots_kernel<NL><<<148, 256>>>(out, cyc, iters, perm, structured);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%d trajectories interleaved in one team, 1 team/SM   %8.1f cycles per team-iteration  %7.2f lane-iterations/us/SM   (%s)\n", NL,
         (double)h / iters, (double)NL * iters / (ms * 1e3), cudaGetErrorString(cudaGetLastError()));
}

// raw TMEM read throughput: W warps of one CTA each stream x32 loads
__global__ void __launch_bounds__(512, 1) ldtm_kernel(long long *cyc, int reps, unsigned *sink) {
  __shared__ uint32_t tbase_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = tbase_s + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(32 * (warp >> 2));
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < reps; ++i) {
    uint32_t r[32];
    tmem_ld32(taddr, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc ^= r[0] ^ r[31];
  }
  const long long t1 = clock64();
  if (acc == 0x12345u) *sink = acc;
  if (threadIdx.x == 0) *cyc = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase_s));
}

template <int TMEM, int CTAS>
static void run(const char *name, double *out, long long *cyc, const int *perm, int structured) {
  const int iters = 4000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  iter_kernel<TMEM, CTAS><<<148 * CTAS, 256>>>(out, cyc, iters, perm, structured);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  iter_kernel<TMEM, CTAS><<<148 * CTAS, 256>>>(out, cyc, iters, perm, structured);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, iter_kernel<TMEM, CTAS>, 256, 0);
  printf("%-44s resident %d/SM  %8.1f cycles/iteration/team  %7.2f lane-iterations/us/SM   (%s)\n", name, nb, (double)h / iters,
         (double)CTAS * iters / (ms * 1e3), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  double *out; long long *cyc, h; int *perm; unsigned *sink;
  cudaMalloc(&out, 1 << 16); cudaMalloc(&cyc, 8); cudaMalloc(&perm, 4096); cudaMalloc(&sink, 4);
  int hp[1024];
  unsigned s = 12345;
  for (int i = 0; i < 1024; ++i) { s = s * 1664525u + 1013904223u; hp[i] = (s >> 8) % 1000; }
  cudaMemcpy(perm, hp, 4096, cudaMemcpyHostToDevice);
  for (int w : {1, 4, 8, 16}) {
    const int reps = 20000;
    ldtm_kernel<<<1, 32 * w>>>(cyc, reps, sink); cudaDeviceSynchronize();
    ldtm_kernel<<<1, 32 * w>>>(cyc, reps, sink); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDTM 32x32b.x32 + wait, %2d warps: %6.1f cycles per load per warp, %7.1f B/cycle/SM   (%s)\n", w, (double)h / reps,
           (double)w * 4096.0 * reps / h, cudaGetErrorString(cudaGetLastError()));
  }
  for (int st = 0; st < 2; ++st) {
    printf("---- gather offsets: %s\n", st ? "structured (few bank conflicts, like the real tables)" : "random (heavy bank conflicts)");
    run<0, 1>("S in registers, 1 team/SM", out, cyc, perm, st);
    run<0, 2>("S in registers, 2 teams/SM", out, cyc, perm, st);
    run<32, 2>("S in TMEM(32)+smem(10), 2 teams/SM", out, cyc, perm, st);
    run<32, 3>("S in TMEM(32)+smem(10), 3 teams/SM", out, cyc, perm, st);
    run<42, 1>("S all in TMEM, 1 team/SM", out, cyc, perm, st);
    run<42, 2>("S all in TMEM, 2 teams/SM", out, cyc, perm, st);
    run_slots<1>(out, cyc, perm, st);
    run_slots<2>(out, cyc, perm, st);
    run_slots<3>(out, cyc, perm, st);
  }
  return 0;
}
