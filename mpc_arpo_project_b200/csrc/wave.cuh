// ADMM "wave" kernel: the multi-RHS solver block for LARGE batches of the n = 81 family (Nx = 10, Nc = Nb = 5; BASELINE
// configs 2 / 3).  Same round contract as admm_block_kernel / admm_tile_kernel: one check_termination period (25
// iterations) of OSQP's ADMM for every live lane of this round, then OSQP's update_info / check_termination /
// is_primal_infeasible / adapt_rho (reference src/trajectorySimulate.py:296 -> prob.solve()).
//
// Linear solve = two dense GEMMs on the FP64 tensor cores.  With the spectral tables of problem.py,
//     x~ = V diag(1/(1+rho lam)) V' rhs        (one V per velocity-sign variant, every lane its own rho),
// a warp owns a tile of 8 lanes and computes   T = R Vp  ->  W = T .* d(rho)  ->  X~ = W Vp'   with mma.sync.m8n8k4.f64:
// the lanes are the M dimension, Vp (81 x 81, 59 KB) sits in shared memory ONCE per CTA and feeds all 8 warps as the
// B operand, and the C fragment of the first GEMM is the A fragment of the second (spectral index ordered
// J(t, c, e) = 8t + 4e + c), so T never leaves the registers.  tools/ubench_dmma.cu: this chain alone runs at 93 % of
// the DMMA pipe with 8 warps per SM.
//
// Everything around the GEMMs is written in UNSCALED variables.  OSQP iterates on the Ruiz-scaled problem
// (x_s = x / D, z_s = E z, y_s = y / E up to the cost scale); substituting, its iteration is an ADMM on the original
// data with a per-row penalty kap_i = rho_i E_i^2 and a per-variable proximal weight sigma / D_j^2:
//     r  = (sigma / D^2) x - c q + A' (kap z - y)          x~ = Vp diag(1/(1+rho lam)) Vp' r,  Vp = D V
//     z~ = A x~ ;  x+ = a x~ + (1-a) x ;  zr = a z~ + (1-a) z ;  z+ = clip(zr + y / kap, l, u) ;  y+ = y + kap (zr - z+)
// (same iterates as oracle/batched_ref.py to 2e-15 relative after 75 iterations).  The point: the unscaled A has the
// SAME few coefficients in every stage (Ad, Ad - Bd K, Bd, C, V_ecr, +-1), so the sparse products are straight-line FMAs
// with constant-bank operands -- no index tables, no per-entry loads.
//
// Thread map.  Thread (g, c) of a warp: g = lane-in-tile, c = member of the lane's quad.  Quad member c owns the horizon
// stages k = 4r + c (slot r = 0..2): the state x_k, the input u_{k-1} that drives it, the slack s_k, and the rows that
// live with them (dynamics row block k, LOS block k, the boxes of u_{k-1} and s_k); d and its pin rows are carried by all
// four.  Only x~_{k-1} and the dual of dynamics block k+1 cross threads (through the tile's shared-memory rows).
// z and y of the thread's 43 rows live in TENSOR MEMORY (tcgen05.ld / st, 248 columns of the thread's own TMEM lane;
// used as per-thread storage, there is no f64 tcgen05.mma), the iterate x in registers.
#pragma once
#include "common.cuh"
#include "team.cuh"       // tensor-memory load / store wrappers

#define WAVE_NX 10
#define WAVE_NC 5
#define WAVE_NB 5
#define WAVE_N 81
#define WAVE_M 136
#define WAVE_KS 21        // k-steps of 4 over the 84 GEMM positions
#define WAVE_NT 11        // 8-wide tiles
#define WAVE_LD 84        // stride of Vp rows and of the tile's n-vector rows: 4 (mod 16) keeps both fragment patterns conflict free
#define WAVE_LDV 52       // stride of the dynamics-dual exchange rows (48 used)
#define WAVE_WARPS 8      // warps (8-lane tiles) per CTA at large batches; a 4-warp instantiation serves batches that would
                          // otherwise leave SMs without a CTA (admm_wave_kernel<4>: 78 % of the 8-warp throughput per SM)
#define WAVE_NVS 28       // variable slots per thread: x 3x4, u 2x2, s 2x5, d 2
#define WAVE_NRS 43       // row slots per thread: dyn 3x4, los 3x5, box-u 2x2, box-s 2x5, pin 2
#define WAVE_INF 1e30

struct WaveHdr {          // byte offsets into the per-variant blob (global -> shared once per CTA)
  int off_Vp;             // [88][WAVE_LD] doubles: Vp[p][J] = D[var(p)] V[var(p)][J], zero rows at unused positions
  int off_lam;            // [88]
  int off_sgD, off_Dv, off_Dinv;          // [NVS][4]: sigma / D^2, D, 1 / D of the variable (slot, quad member)
  int off_ka, off_kb, off_kie;            // [NRS][4]: kap_i = ka rho + kb (free rows: ka = 0, kb = rho_min E^2; else ka = class factor x E^2, kb = 0); 1 / ka
  int off_Ev, off_Einv;   // [NRS][4]
  int off_M1;             // [2][16]: Ad, Ad - Bd K (slot 1 mixes the two kinds of stage)
  int total;
};

struct WaveConst {        // unscaled problem data, passed by value (constant bank operands)
  double Ad[16], Acl[16], Bd[8], C[20], Vecr[5];
  double Q[16], QN[16], Ru[4], Rs[5];
  double qx[4], qN[4];    // c * q of a stage k < Nx / k = Nx
  double ulim[2], r_p;
  double c, cinv, sigma, alpha, eps_abs, eps_rel, eps_pinf, adapt_tol, qn_unscaled, qn_scaled;
  int check_every, adaptive, adapt_interval, max_iter;
};

struct WaveArgs {
  WaveHdr hdr;
  WaveConst k;
  const unsigned char *blob[4];
  int B;
  const int *cnt;        // [4] lanes per variant this round
  const int *list;       // [4][B]
  double *xs, *zs, *ys;  // [B][n], [B][m], [B][m] OSQP-scaled iterates (the format every solver block shares)
  double *rho;
  int *iter, *status;
  const double *par;     // [7][B]
  double *u0;            // [2][B]
  uint8_t *lane_state;
  int *flip;
  unsigned long long *iter_total;
};

__device__ __forceinline__ void wave_dmma(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double wq_max(double v) {
  v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmax(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ double wq_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// ---- tensor memory <-> doubles (thread-private columns; 2 columns per double)
__device__ __forceinline__ void wtm_ld4(uint32_t a, double (&v)[4]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a));
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) :: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = u2d(r[2 * i], r[2 * i + 1]);
}
__device__ __forceinline__ void wtm_ld2(uint32_t a, double (&v)[2]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]) :: "memory");
  v[0] = u2d(r[0], r[1]);
  v[1] = u2d(r[2], r[3]);
}
__device__ __forceinline__ void wtm_st4(uint32_t a, const double (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a),
               "r"((uint32_t)__double2loint(v[0])), "r"((uint32_t)__double2hiint(v[0])), "r"((uint32_t)__double2loint(v[1])),
               "r"((uint32_t)__double2hiint(v[1])), "r"((uint32_t)__double2loint(v[2])), "r"((uint32_t)__double2hiint(v[2])),
               "r"((uint32_t)__double2loint(v[3])), "r"((uint32_t)__double2hiint(v[3])) : "memory");
}
__device__ __forceinline__ void wtm_st2(uint32_t a, const double (&v)[2]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"((uint32_t)__double2loint(v[0])),
               "r"((uint32_t)__double2hiint(v[0])), "r"((uint32_t)__double2loint(v[1])), "r"((uint32_t)__double2hiint(v[1])) : "memory");
}

// tensor-memory columns of a thread (32-bit columns; see the header comment for the row groups)
#define WTM_ZD(r) (36 * (r))            // z dyn[4]
#define WTM_YD(r) (36 * (r) + 8)        // y dyn[4]
#define WTM_ZS(r) (36 * (r) + 16)       // z los[0..3]
#define WTM_ZS4(r) (36 * (r) + 24)      // z los[4], pad
#define WTM_YS(r) (36 * (r) + 28)       // y los[0..3]   (los[4] is a free row: y = 0)
#define WTM_ZU(r) (108 + 28 * (r))      // z box-u[2]
#define WTM_YU(r) (108 + 28 * (r) + 4)  // y box-u[2]
#define WTM_ZB(r) (108 + 28 * (r) + 8)  // z box-s[0..3]
#define WTM_ZYB4(r) (108 + 28 * (r) + 16)   // z box-s[4], y box-s[4]
#define WTM_YB(r) (108 + 28 * (r) + 20) // y box-s[0..3]
#define WTM_PIN 164                      // z pin[2], y pin[2]
#define WTM_DD(r) (172 + 8 * (r))        // projected delta-y of the block's last iteration: dyn[4]
#define WTM_DS(r) (196 + 8 * (r))        // los[0..3], r < 2
#define WTM_DU(r) (212 + 16 * (r))       // box-u[2]
#define WTM_DB(r) (212 + 16 * (r) + 4)   // box-s[0..3]
#define WTM_DB4(r) (212 + 16 * (r) + 12) // box-s[4], pad
#define WTM_DP 244                       // pin[2]        -> 248 columns

// slot numbering of the [slot][4] tables
#define WVS_X(r, j) ((r) * 4 + (j))
#define WVS_U(r, j) (12 + (r) * 2 + (j))
#define WVS_S(r, j) (16 + (r) * 5 + (j))
#define WVS_D(j) (26 + (j))
#define WRS_DYN(r, i) ((r) * 4 + (i))
#define WRS_LOS(r, i) (12 + (r) * 5 + (i))
#define WRS_BU(r, i) (27 + (r) * 2 + (i))
#define WRS_BS(r, i) (31 + (r) * 5 + (i))
#define WRS_PIN(i) (41 + (i))

// ---- batched tensor-memory loads: issue, issue, ..., ONE wait that names every destination register (the compiler must
//      not move a use of them above the wait)
__device__ __forceinline__ void wtm_ld8_nw(uint32_t a, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a));
}
__device__ __forceinline__ void wtm_ld4_nw(uint32_t a, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
#define WTM_R8(a) "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7])
#define WTM_R4(a) "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3])
__device__ __forceinline__ void wtm_wait_stage(uint32_t (&a)[8], uint32_t (&b)[8], uint32_t (&c)[8], uint32_t (&d)[4], uint32_t (&e)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : WTM_R8(a), WTM_R8(b), WTM_R8(c), WTM_R4(d), WTM_R8(e) :: "memory");
}
__device__ __forceinline__ void wtm_wait_box(uint32_t (&a)[4], uint32_t (&b)[4], uint32_t (&c)[8], uint32_t (&d)[4], uint32_t (&e)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : WTM_R4(a), WTM_R4(b), WTM_R8(c), WTM_R4(d), WTM_R8(e) :: "memory");
}
__device__ __forceinline__ void wtm_wait_dy(uint32_t (&a)[8], uint32_t (&b)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : WTM_R8(a), WTM_R8(b) :: "memory");
}
__device__ __forceinline__ void wtm_wait_dybox(uint32_t (&a)[4], uint32_t (&b)[8], uint32_t (&c)[4]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : WTM_R4(a), WTM_R8(b), WTM_R4(c) :: "memory");
}
__device__ __forceinline__ double wd(const uint32_t *r, int i) { return u2d(r[2 * i], r[2 * i + 1]); }
__device__ __forceinline__ void wtm_st8_raw(uint32_t a, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

// spectral index of column n of an 8-wide tile: with this order BOTH GEMMs read their B fragments from the one table Vp
// (row stride 84 = 4 mod 16) without bank conflicts, and the C fragment of GEMM 1 is still the A fragment of GEMM 2
__device__ __forceinline__ int wave_pi(int n) { return n < 4 ? n : (n ^ 1); }        // {0,1,2,3,5,4,7,6}

// ADMM row updates (OSQP update_z / update_y in the unscaled variables), one flavour per kind of row; each returns
// v = kap z - y, the row's entry of the next A' product.
struct WaveRow {
  double alpha, oma, rho, rinv;
  const double *ka, *kb, *kie;      // + 4 * slot (+ c folded in)
  // equality row (l = u = b): the projection is b itself
  __device__ __forceinline__ double eq(int slot, double zt, double &z, double &y, double b, double &dy) const {
    const double k = rho * ka[4 * slot];
    const double zr = alpha * zt + oma * z;
    dy = k * (zr - b);
    y += dy;
    z = b;
    return k * b - y;
  }
  // free row (both bounds infinite): kap = rho_min E^2, y stays 0
  __device__ __forceinline__ double fr(int slot, double zt, double &z) const {
    z = alpha * zt + oma * z;
    return kb[4 * slot] * z;
  }
  // general row; lo / hi may be -+WAVE_INF.  Free rows pass through here too when the kind is only known per thread
  // (ka = 0, kb = rho_min E^2, y = 0, infinite bounds).
  __device__ __forceinline__ double gen(int slot, double zt, double &z, double &y, double lo, double hi, double &dy) const {
    const double k = fma(ka[4 * slot], rho, kb[4 * slot]), ik = rinv * kie[4 * slot];
    const double zr = alpha * zt + oma * z;
    const double zn = fmin(fmax(zr + ik * y, lo), hi);
    dy = k * (zr - zn);
    y += dy;
    z = zn;
    return k * zn - y;
  }
  // lower bound only
  __device__ __forceinline__ double lo1(int slot, double zt, double &z, double &y, double lo, double &dy) const {
    const double k = rho * ka[4 * slot], ik = rinv * kie[4 * slot];
    const double zr = alpha * zt + oma * z;
    const double zn = fmax(zr + ik * y, lo);
    dy = k * (zr - zn);
    y += dy;
    z = zn;
    return k * zn - y;
  }
  __device__ __forceinline__ double vee(int slot, double z, double y) const { return fma(ka[4 * slot], rho, kb[4 * slot]) * z - y; }
};

template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 1) admm_wave_kernel(const __grid_constant__ WaveArgs a) {
  constexpr int NX = WAVE_NX, NC = WAVE_NC, NB = WAVE_NB, N = WAVE_N, M = WAVE_M, KS = WAVE_KS, NT = WAVE_NT, LD = WAVE_LD,
                LDV = WAVE_LDV;
  constexpr int nX = 4 * (NX + 1), RB = nX + 5 * (NX + 1);       // first input/slack variable; first box row
  extern __shared__ __align__(128) unsigned char smem[];
  const WaveHdr &h = a.hdr;
  const WaveConst &K = a.k;
  const int warp = threadIdx.x >> 5, lid = threadIdx.x & 31, g = lid >> 2, c = lid & 3;

  // ---- which (variant, group of 8 * WARPS lanes) is this CTA
  int b = blockIdx.x, v = 0, cnt_v = 0;
  for (; v < 4; ++v) {
    cnt_v = a.cnt[v];
    const int nt = (cnt_v + 8 * WARPS - 1) / (8 * WARPS);
    if (b < nt) break;
    b -= nt;
  }
  if (v == 4) return;
  for (int o = threadIdx.x * 16; o < h.total; o += 32 * WARPS * 16)
    *reinterpret_cast<int4 *>(smem + o) = *reinterpret_cast<const int4 *>(a.blob[v] + o);
  __shared__ uint32_t s_tmem;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tmem)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = s_tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(256 * (warp >> 2));

  const double *Vp = reinterpret_cast<const double *>(smem + h.off_Vp);
  const double *lam = reinterpret_cast<const double *>(smem + h.off_lam);
  const double *sgD = reinterpret_cast<const double *>(smem + h.off_sgD) + c;
  const double *Dv = reinterpret_cast<const double *>(smem + h.off_Dv) + c;
  const double *Dinv = reinterpret_cast<const double *>(smem + h.off_Dinv) + c;
  const double *kat = reinterpret_cast<const double *>(smem + h.off_ka) + c;
  const double *kbt = reinterpret_cast<const double *>(smem + h.off_kb) + c;
  const double *kiet = reinterpret_cast<const double *>(smem + h.off_kie) + c;
  const double *Ev = reinterpret_cast<const double *>(smem + h.off_Ev) + c;
  const double *Einv = reinterpret_cast<const double *>(smem + h.off_Einv) + c;
  const double *M1tab = reinterpret_cast<const double *>(smem + h.off_M1);

  // per-warp tile rows: n-vector (r, then x~), dynamics-dual exchange, spectral weights
  double *wbase = reinterpret_cast<double *>(smem + h.total) + (size_t)warp * 8 * (LD + LDV + LD);
  double *nb = wbase + g * LD;
  double *vb = wbase + 8 * LD + g * LDV;
  double *dscb = wbase + 8 * (LD + LDV) + g * LD;

  const int pos = (b * WARPS + warp) * 8 + g;
  const bool valid = pos < cnt_v;
  const bool tile_live = __ballot_sync(0xffffffffu, valid) != 0;
  const int ln = valid ? a.list[(size_t)v * a.B + pos] : 0;
  const size_t Bz = a.B;

  // velocity signs of this variant (simhelpers.py:66-67: C1 = sign(vx^), C2 = sign(vy^), sign(0) = +1)
  const double c1 = (v & 1) ? -1.0 : 1.0, c2 = (v & 2) ? -1.0 : 1.0;

  // stage bookkeeping of this quad member: slot r <-> stage k = 4r + c
  bool st_ok[3], st_u[2], st_s[2], st_b[3];
  double ge1[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int k = 4 * r + c;
    st_ok[r] = k <= NX;
    st_b[r] = k <= NB;
    ge1[r] = (k >= 1 && k <= NX) ? 1.0 : 0.0;
    if (r < 2) {
      st_u[r] = k >= 1 && k <= NC;
      st_s[r] = k < NC;
    }
  }
  // slot 1 mixes stages that propagate with Ad (k <= Nc) and with Ad - Bd K: matrix of stage k (rows) and of stage k+1 (A')
  const double *M1row = M1tab + ((4 + c) <= NC ? 0 : 16);
  const double *M1nxt = M1tab + ((5 + c) <= NC ? 0 : 16);

  // GEMM positions of this thread's variables
  auto px = [&](int r, int j) { return 16 * r + 4 * j + c; };
  auto pu = [&](int r, int j) { return r == 0 ? 48 + 3 * j + (c - 1) : 54 + 2 * j + c; };
  auto ps = [&](int r, int j) { return r == 0 ? 58 + 4 * j + c : 78 + j; };
  auto pd = [&](int j) { return 35 + 4 * j; };

  // ---- lane parameters and iterates
  double prm[7];
#pragma unroll
  for (int q = 0; q < 7; ++q) prm[q] = valid ? a.par[q * Bz + ln] : 0.0;
  double rho = valid ? a.rho[ln] : 1.0;
  int iter = valid ? a.iter[ln] : 0;
  double xk[3][4], uk[2][2], sk[2][5], dd[2];
  int flipf = 0;
  {
    const double *xg = a.xs + (size_t)ln * N;
    const double *zg = a.zs + (size_t)ln * M, *yg = a.ys + (size_t)ln * M;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int k = 4 * r + c;
      const bool ok = valid && st_ok[r];
      double zd[4], yd[4], zs4[4], ys4[4], zs5[2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xk[r][j] = ok ? xg[4 * k + j] * Dv[4 * WVS_X(r, j)] : 0.0;
        zd[j] = ok ? zg[4 * k + j] * Einv[4 * WRS_DYN(r, j)] : 0.0;
        yd[j] = ok ? yg[4 * k + j] * Ev[4 * WRS_DYN(r, j)] : 0.0;
        zs4[j] = ok ? zg[nX + 5 * k + j] * Einv[4 * WRS_LOS(r, j)] : 0.0;
        ys4[j] = ok ? yg[nX + 5 * k + j] * Ev[4 * WRS_LOS(r, j)] : 0.0;
      }
      zs5[0] = ok ? zg[nX + 5 * k + 4] * Einv[4 * WRS_LOS(r, 4)] : 0.0;
      zs5[1] = 0.0;
      wtm_st4(tb + WTM_ZD(r), zd);
      wtm_st4(tb + WTM_YD(r), yd);
      wtm_st4(tb + WTM_ZS(r), zs4);
      wtm_st2(tb + WTM_ZS4(r), zs5);
      wtm_st4(tb + WTM_YS(r), ys4);
      if (ok && st_b[r] && prm[4] < MPCB_RHO_TOL * Einv[4 * WRS_LOS(r, 3)]) flipf = 1;      // E * val - 0 < RHO_TOL: OSQP would re-type the row
      if (r < 2) {
        const bool oku = valid && st_u[r], oks = valid && st_s[r];
        double zu[2], yu[2], zb[4], yb[4], zyb[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uk[r][j] = oku ? xg[nX + 7 * (k - 1) + j] * Dv[4 * WVS_U(r, j)] : 0.0;
          zu[j] = oku ? zg[RB + 7 * (k - 1) + j] * Einv[4 * WRS_BU(r, j)] : 0.0;
          yu[j] = oku ? yg[RB + 7 * (k - 1) + j] * Ev[4 * WRS_BU(r, j)] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) sk[r][j] = oks ? xg[nX + 7 * k + 2 + j] * Dv[4 * WVS_S(r, j)] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          zb[j] = oks ? zg[RB + 7 * k + 2 + j] * Einv[4 * WRS_BS(r, j)] : 0.0;
          yb[j] = oks ? yg[RB + 7 * k + 2 + j] * Ev[4 * WRS_BS(r, j)] : 0.0;
        }
        zyb[0] = oks ? zg[RB + 7 * k + 6] * Einv[4 * WRS_BS(r, 4)] : 0.0;
        zyb[1] = oks ? yg[RB + 7 * k + 6] * Ev[4 * WRS_BS(r, 4)] : 0.0;
        wtm_st2(tb + WTM_ZU(r), zu);
        wtm_st2(tb + WTM_YU(r), yu);
        wtm_st4(tb + WTM_ZB(r), zb);
        wtm_st2(tb + WTM_ZYB4(r), zyb);
        wtm_st4(tb + WTM_YB(r), yb);
      }
    }
    double pin4[4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      dd[j] = valid ? xg[N - 2 + j] * Dv[4 * WVS_D(j)] : 0.0;
      pin4[j] = valid ? zg[M - 2 + j] * Einv[4 * WRS_PIN(j)] : 0.0;
      pin4[2 + j] = valid ? yg[M - 2 + j] * Ev[4 * WRS_PIN(j)] : 0.0;
    }
    wtm_st4(tb + WTM_PIN, pin4);
    tmem_wait_st();
  }
  flipf = __shfl_xor_sync(0xffffffffu, flipf, 1) | flipf;
  flipf = __shfl_xor_sync(0xffffffffu, flipf, 2) | flipf;
  if (valid && flipf && c == 0) a.flip[ln] = 1;

  auto set_dsc = [&]() {                 // spectral weights of this lane: thread (g, c) fills J = c (mod 4)
#pragma unroll
    for (int q = 0; q < LD / 4; ++q) {
      const int J = 4 * q + c;
      dscb[J] = J < N ? 1.0 / (1.0 + rho * lam[J]) : 0.0;
    }
  };
  set_dsc();
  for (int q = lid; q < 8 * LD; q += 32) wbase[q] = 0.0;            // n-vector rows: unused positions must stay finite
  for (int q = lid; q < 8 * LDV; q += 32) wbase[8 * LD + q] = 0.0;
  __syncwarp();

  const double alpha = K.alpha, oma = 1.0 - K.alpha;
  WaveRow R;
  R.alpha = alpha; R.oma = oma; R.rho = rho; R.rinv = 1.0 / rho; R.ka = kat; R.kb = kbt; R.kie = kiet;

  // ---- bounds of this step (unscaled): only x^ (dynamics block 0), the velocity 1-norm bound and the disturbance pin move
  auto los_lo = [&](int r, int i) { return st_b[r] ? (i == 0 || i == 1 ? 1.0 : (i == 2 ? K.r_p : (i == 3 ? 0.0 : -WAVE_INF))) : -WAVE_INF; };
  auto los_hi = [&](int r, int i) { return (st_b[r] && i == 3) ? prm[4] : WAVE_INF; };

  // r accumulators of the next linear solve (variables this thread owns)
  double rx[3][4], ru[2][2], rs[2][5], rd[2], dsum[2];

  // A' contributions of one stage's rows to the thread's own variables; dynamics duals are also published for the neighbour
  auto at_stage = [&](int r, const double (&vd)[4], const double (&vs)[5]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double acc = rx[r][j] - vd[j];
#pragma unroll
      for (int i = 0; i < 3; ++i) acc = fma(K.C[4 * i + j], vs[i], acc);
      if (j == 2) acc = fma(c1, vs[3], acc);
      if (j == 3) acc = fma(c2, vs[3], acc);
      if (j == 1) acc += vs[4];                        // row 4 of the LOS block = [0 1 0 0] (no debris: slope = 0)
      rx[r][j] = acc;
      vb[16 * r + 4 * j + c] = vd[j];
    }
    dsum[0] = fma(ge1[r], vd[0], dsum[0]);
    dsum[1] = fma(ge1[r], vd[1], dsum[1]);
  };
  auto at_box = [&](int r, const double (&vd)[4], const double (&vs)[5], const double (&vu)[2], const double (&vbs)[5]) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double acc = ru[r][j] + vu[j];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc = fma(K.Bd[2 * i + j], vd[i], acc);
      ru[r][j] = acc;
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) rs[r][i] = fma(K.Vecr[i], vs[i], rs[r][i] + vbs[i]);
  };
  // second half of A': the dynamics duals of stage k+1 (a neighbour's) through that stage's propagation matrix; d; store r
  auto at_finish = [&](const double (&vpin)[2]) {
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int kn = 4 * r + c + 1;                    // stage k+1 lives with quad member (c+1) & 3, slot r (+1 when c = 3)
      const bool okn = kn <= NX;
      const int off = (c == 3) ? 16 * (r + 1) : 16 * r + (c + 1);
      double vn[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) vn[i] = okn ? vb[off + 4 * i] : 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double acc = rx[r][j];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double mij = (r == 0) ? K.Ad[4 * i + j] : (r == 2 ? K.Acl[4 * i + j] : M1nxt[4 * i + j]);
          acc = fma(mij, vn[i], acc);
        }
        rx[r][j] = acc;
      }
    }
    rd[0] += wq_sum(dsum[0]) + vpin[0];
    rd[1] += wq_sum(dsum[1]) + vpin[1];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (st_ok[r]) nb[px(r, j)] = rx[r][j];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (st_u[r]) nb[pu(r, j)] = ru[r][j];
#pragma unroll
      for (int j = 0; j < 5; ++j)
        if (st_s[r]) nb[ps(r, j)] = rs[r][j];
    }
    if (c == 3) {
      nb[pd(0)] = rd[0];
      nb[pd(1)] = rd[1];
    }
    __syncwarp();
  };
  auto r_init = [&]() {                  // r = (sigma / D^2) x - c q
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double q = (r == 2 && c == 2) ? K.qN[j] : K.qx[j];
        rx[r][j] = sgD[4 * WVS_X(r, j)] * xk[r][j] - q;
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int j = 0; j < 2; ++j) ru[r][j] = sgD[4 * WVS_U(r, j)] * uk[r][j];
#pragma unroll
      for (int j = 0; j < 5; ++j) rs[r][j] = sgD[4 * WVS_S(r, j)] * sk[r][j];
    }
    rd[0] = sgD[4 * WVS_D(0)] * dd[0];
    rd[1] = sgD[4 * WVS_D(1)] * dd[1];
    dsum[0] = dsum[1] = 0.0;
  };

  // ---- r of the round's first iteration from the loaded (z, y)
  {
    r_init();
    double vpin[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double okf = st_ok[r] ? 1.0 : 0.0;
      double zd[4], yd[4], zs4[4], ys4[4], zs5[2], vd[4], vs[5];
      wtm_ld4(tb + WTM_ZD(r), zd);
      wtm_ld4(tb + WTM_YD(r), yd);
      wtm_ld4(tb + WTM_ZS(r), zs4);
      wtm_ld2(tb + WTM_ZS4(r), zs5);
      wtm_ld4(tb + WTM_YS(r), ys4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        vd[i] = R.vee(WRS_DYN(r, i), zd[i], yd[i]);
        vs[i] = R.vee(WRS_LOS(r, i), zs4[i], ys4[i]);
      }
      vs[4] = R.vee(WRS_LOS(r, 4), zs5[0], 0.0);
      at_stage(r, vd, vs);
      if (r < 2) {
        double zu[2], yu[2], zb[4], yb[4], zyb[2], vu[2], vbs[5];
        wtm_ld2(tb + WTM_ZU(r), zu);
        wtm_ld2(tb + WTM_YU(r), yu);
        wtm_ld4(tb + WTM_ZB(r), zb);
        wtm_ld2(tb + WTM_ZYB4(r), zyb);
        wtm_ld4(tb + WTM_YB(r), yb);
        const double uf = st_u[r] ? 1.0 : 0.0, sf = st_s[r] ? 1.0 : 0.0;
#pragma unroll
        for (int i = 0; i < 2; ++i) vu[i] = R.vee(WRS_BU(r, i), zu[i], yu[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) vbs[i] = R.vee(WRS_BS(r, i), zb[i], yb[i]);
        vbs[4] = R.vee(WRS_BS(r, 4), zyb[0], zyb[1]);
        at_box(r, vd, vs, vu, vbs);
      }
    }
    double pin4[4];
    wtm_ld4(tb + WTM_PIN, pin4);
    vpin[0] = R.vee(WRS_PIN(0), pin4[0], pin4[2]);
    vpin[1] = R.vee(WRS_PIN(1), pin4[1], pin4[3]);
    at_finish(vpin);
  }

  // =============================== check_every ADMM iterations ===============================
  const int pi0 = wave_pi(2 * c), pi1 = wave_pi(2 * c + 1);          // spectral offsets of this thread's two C-fragment columns
  for (int it = 0; tile_live && it < K.check_every; ++it) {
    const bool last = it == K.check_every - 1;
    // ---- x~ = Vp diag(d) Vp' r on the tensor cores
    {
      double af[KS];
#pragma unroll
      for (int s = 0; s < KS; ++s) af[s] = nb[4 * s + c];
      __syncwarp();
      double acc[2 * NT];
#pragma unroll
      for (int j = 0; j < 2 * NT; ++j) acc[j] = 0.0;
      const double *vb1 = Vp + c * LD + wave_pi(g);                  // B[k = 4s + c][J = 8t + pi(g)]
#pragma unroll
      for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int t = 0; t < NT; ++t) wave_dmma(acc[2 * t], acc[2 * t + 1], af[s], vb1[4 * s * LD + 8 * t]);
      // thread (g, c) holds T[lane g][J = 8t + pi(2c + e)] in acc[2t + e]; J >= 84 (last tile, c >= 2) is padding: weight 0
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const bool pad = (t == NT - 1) && c >= 2;
        acc[2 * t] *= pad ? 0.0 : dscb[8 * t + pi0];
        acc[2 * t + 1] *= pad ? 0.0 : dscb[8 * t + pi1];
      }
      // k-step (2t + e) of GEMM 2 covers J = 8t + pi(2c' + e), c' = 0..3: B[J][p = 8t' + g] = Vp[8t' + g][8t + pi(2c + e)]
      const double *vb2 = Vp + g * LD;
#pragma unroll
      for (int tp = 0; tp < NT; ++tp) {
        // two accumulator pairs per output tile: a lone warp is otherwise bound by the latency of a 22-long DMMA chain
        double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          wave_dmma(x0, x1, acc[2 * t], vb2[8 * tp * LD + 8 * t + pi0]);
          wave_dmma(y0, y1, acc[2 * t + 1], vb2[8 * tp * LD + 8 * t + pi1]);
        }
        const int p = 8 * tp + 2 * c;
        if (p < LD) *reinterpret_cast<double2 *>(nb + p) = make_double2(x0 + y0, x1 + y1);
      }
      __syncwarp();
    }
    // ---- rows: z~ = A x~, x / z / y updates, and the A' product of the NEXT iteration accumulated row by row
    double dt[2];
    dt[0] = nb[pd(0)];
    dt[1] = nb[pd(1)];
    dd[0] = alpha * dt[0] + oma * dd[0];
    dd[1] = alpha * dt[1] + oma * dd[1];
    rd[0] = sgD[4 * WVS_D(0)] * dd[0];
    rd[1] = sgD[4 * WVS_D(1)] * dd[1];
    dsum[0] = dsum[1] = 0.0;
    double vpin[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      // tensor-memory loads of the slot's rows first: they complete while the products below are computed
      uint32_t qzd[8], qyd[8], qzs[8], qz5[4], qys[8];
      wtm_ld8_nw(tb + WTM_ZD(r), qzd);
      wtm_ld8_nw(tb + WTM_YD(r), qyd);
      wtm_ld8_nw(tb + WTM_ZS(r), qzs);
      wtm_ld4_nw(tb + WTM_ZS4(r), qz5);
      wtm_ld8_nw(tb + WTM_YS(r), qys);
      // x~ of this stage (own) and of stage k-1 (quad member (c-1) & 3, slot r; slot r-1 when c = 0)
      double xt[4], xm[4], ut[2] = {0.0, 0.0}, sl[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
      {
        const bool okm = ge1[r] != 0.0;
        const int off = (c == 0) ? 16 * (r - 1) + 3 : 16 * r + (c - 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          xt[j] = st_ok[r] ? nb[px(r, j)] : 0.0;
          xm[j] = okm ? nb[off + 4 * j] : 0.0;
          xk[r][j] = alpha * xt[j] + oma * xk[r][j];
          const double q = (r == 2 && c == 2) ? K.qN[j] : K.qx[j];
          rx[r][j] = sgD[4 * WVS_X(r, j)] * xk[r][j] - q;
        }
        if (r < 2) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            ut[j] = st_u[r] ? nb[pu(r, j)] : 0.0;
            uk[r][j] = alpha * ut[j] + oma * uk[r][j];
            ru[r][j] = sgD[4 * WVS_U(r, j)] * uk[r][j];
          }
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            sl[j] = st_s[r] ? nb[ps(r, j)] : 0.0;
            sk[r][j] = alpha * sl[j] + oma * sk[r][j];
            rs[r][j] = sgD[4 * WVS_S(r, j)] * sk[r][j];
          }
        }
      }
      double zt_d[4], zt_s[5];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double acc = -xt[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double mij = (r == 0) ? K.Ad[4 * i + j] : (r == 2 ? K.Acl[4 * i + j] : M1row[4 * i + j]);
          acc = fma(mij, xm[j], acc);
        }
        if (r < 2) {
          acc = fma(K.Bd[2 * i], ut[0], acc);
          acc = fma(K.Bd[2 * i + 1], ut[1], acc);
        }
        if (i < 2) acc = fma(ge1[r], dt[i], acc);
        zt_d[i] = acc;
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc = fma(K.C[4 * i + j], xt[j], acc);
        zt_s[i] = acc;
      }
      zt_s[3] = c1 * xt[2] + c2 * xt[3];
      zt_s[4] = xt[1];
      if (r < 2) {
#pragma unroll
        for (int i = 0; i < 5; ++i) zt_s[i] = fma(K.Vecr[i], sl[i], zt_s[i]);
      }
      wtm_wait_stage(qzd, qyd, qzs, qz5, qys);
      double zd[4], yd[4], zs4[4], ys4[4], zs5[2], vd[4], vs[5], dyd[4], dys[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        zd[i] = wd(qzd, i); yd[i] = wd(qyd, i); zs4[i] = wd(qzs, i); ys4[i] = wd(qys, i);
      }
      zs5[0] = wd(qz5, 0);
      zs5[1] = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double bd = (r == 0 && c == 0) ? -prm[i] : 0.0;              // dynamics block 0: -x^ ; the others: 0
        vd[i] = R.eq(WRS_DYN(r, i), zt_d[i], zd[i], yd[i], bd, dyd[i]);
      }
      if (r == 0) {                      // stages 0..3 <= Nb: LOS cone / keep-out rows have a lower bound, the 1-norm row two
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          double dl;
          vs[i] = R.lo1(WRS_LOS(r, i), zt_s[i], zs4[i], ys4[i], i == 2 ? K.r_p : 1.0, dl);
          dys[i] = fmin(dl, 0.0);        // no upper bound: only dy <= 0 certifies infeasibility
        }
        vs[3] = R.gen(WRS_LOS(r, 3), zt_s[3], zs4[3], ys4[3], 0.0, prm[4], dys[3]);
      } else if (r == 1) {               // stages 4..7: bounded up to Nb, free beyond (per thread)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          double dl;
          vs[i] = R.gen(WRS_LOS(r, i), zt_s[i], zs4[i], ys4[i], los_lo(r, i), los_hi(r, i), dl);
          dys[i] = st_b[r] ? ((i < 3) ? fmin(dl, 0.0) : dl) : 0.0;
        }
      } else {                           // stages 8..: free rows
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          vs[i] = R.fr(WRS_LOS(r, i), zt_s[i], zs4[i]);
          dys[i] = 0.0;
        }
      }
      vs[4] = R.fr(WRS_LOS(r, 4), zt_s[4], zs5[0]);                        // half-plane row: free without debris
      wtm_st4(tb + WTM_ZD(r), zd);
      wtm_st4(tb + WTM_YD(r), yd);
      wtm_st4(tb + WTM_ZS(r), zs4);
      wtm_st2(tb + WTM_ZS4(r), zs5);
      if (r < 2) wtm_st4(tb + WTM_YS(r), ys4);
      if (last) {
        wtm_st4(tb + WTM_DD(r), dyd);
        if (r < 2) wtm_st4(tb + WTM_DS(r), dys);
      }
      at_stage(r, vd, vs);
      if (r < 2) {
        uint32_t qzu[4], qyu[4], qzb[8], qzy4[4], qyb[8];
        wtm_ld4_nw(tb + WTM_ZU(r), qzu);
        wtm_ld4_nw(tb + WTM_YU(r), qyu);
        wtm_ld8_nw(tb + WTM_ZB(r), qzb);
        wtm_ld4_nw(tb + WTM_ZYB4(r), qzy4);
        wtm_ld8_nw(tb + WTM_YB(r), qyb);
        wtm_wait_box(qzu, qyu, qzb, qzy4, qyb);
        double zu[2], yu[2], zb[4], yb[4], zyb[2], vu[2], vbs[5], dyu[2], dyb[4], dyb4[2];
        zu[0] = wd(qzu, 0); zu[1] = wd(qzu, 1); yu[0] = wd(qyu, 0); yu[1] = wd(qyu, 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          zb[i] = wd(qzb, i); yb[i] = wd(qyb, i);
        }
        zyb[0] = wd(qzy4, 0); zyb[1] = wd(qzy4, 1);
#pragma unroll
        for (int i = 0; i < 2; ++i) vu[i] = R.gen(WRS_BU(r, i), ut[i], zu[i], yu[i], -K.ulim[i], K.ulim[i], dyu[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          double dl;
          vbs[i] = R.lo1(WRS_BS(r, i), sl[i], zb[i], yb[i], 0.0, dl);
          dyb[i] = fmin(dl, 0.0);
        }
        {
          double dl;
          vbs[4] = R.lo1(WRS_BS(r, 4), sl[4], zyb[0], zyb[1], 0.0, dl);
          dyb4[0] = fmin(dl, 0.0);
          dyb4[1] = 0.0;
        }
        wtm_st2(tb + WTM_ZU(r), zu);
        wtm_st2(tb + WTM_YU(r), yu);
        wtm_st4(tb + WTM_ZB(r), zb);
        wtm_st2(tb + WTM_ZYB4(r), zyb);
        wtm_st4(tb + WTM_YB(r), yb);
        if (last) {
          wtm_st2(tb + WTM_DU(r), dyu);
          wtm_st4(tb + WTM_DB(r), dyb);
          wtm_st2(tb + WTM_DB4(r), dyb4);
        }
        at_box(r, vd, vs, vu, vbs);
      }
    }
    {
      double pin4[4], dyp[2];
      wtm_ld4(tb + WTM_PIN, pin4);
      vpin[0] = R.eq(WRS_PIN(0), dt[0], pin4[0], pin4[2], prm[5], dyp[0]);
      vpin[1] = R.eq(WRS_PIN(1), dt[1], pin4[1], pin4[3], prm[6], dyp[1]);
      wtm_st4(tb + WTM_PIN, pin4);
      if (last) wtm_st2(tb + WTM_DP, dyp);
    }
    tmem_wait_st();
    at_finish(vpin);
  }
  iter += K.check_every;

  // =============================== update_info / check_termination (OSQP auxil.c) ===============================
  // nb <- x (positions), so that A x can read the neighbour's x_{k-1}
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (st_ok[r]) nb[px(r, j)] = xk[r][j];
  __syncwarp();
  double mu[8], ms[6], lhs = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) mu[q] = 0.0;
#pragma unroll
  for (int q = 0; q < 6; ++q) ms[q] = 0.0;
  // two A' products (y, and the projected delta-y of the last iteration), accumulated like the iteration's; run one after the
  // other through the same accumulators
  double aty_x[3][4], aty_u[2][2], aty_s[2][5], aty_d[2];
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) rx[r][j] = 0.0;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int j = 0; j < 2; ++j) ru[r][j] = 0.0;
#pragma unroll
      for (int j = 0; j < 5; ++j) rs[r][j] = 0.0;
    }
    rd[0] = rd[1] = dsum[0] = dsum[1] = 0.0;
    double vpin[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double okf = st_ok[r] ? 1.0 : 0.0;
      double vd[4], vs[5];
      if (pass == 0) {
        double zd[4], yd[4], zs4[4], ys4[4], zs5[2], xm[4];
        wtm_ld4(tb + WTM_ZD(r), zd);
        wtm_ld4(tb + WTM_YD(r), yd);
        wtm_ld4(tb + WTM_ZS(r), zs4);
        wtm_ld2(tb + WTM_ZS4(r), zs5);
        wtm_ld4(tb + WTM_YS(r), ys4);
        const bool okm = ge1[r] != 0.0;
        const int off = (c == 0) ? 16 * (r - 1) + 3 : 16 * r + (c - 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) xm[j] = okm ? nb[off + 4 * j] : 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {                 // dynamics rows: A x, residuals
          double ax = -xk[r][i];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const double mij = (r == 0) ? K.Ad[4 * i + j] : (r == 2 ? K.Acl[4 * i + j] : M1row[4 * i + j]);
            ax = fma(mij, xm[j], ax);
          }
          if (r < 2) {
            ax = fma(K.Bd[2 * i], uk[r][0], ax);
            ax = fma(K.Bd[2 * i + 1], uk[r][1], ax);
          }
          if (i < 2) ax = fma(ge1[r], dd[i], ax);
          const double e = Ev[4 * WRS_DYN(r, i)], pv = okf * (ax - zd[i]), zz = okf * zd[i], aa = okf * ax;
          mu[0] = fmax(mu[0], fabs(pv)); mu[1] = fmax(mu[1], fabs(zz)); mu[2] = fmax(mu[2], fabs(aa));
          ms[0] = fmax(ms[0], fabs(e * pv)); ms[1] = fmax(ms[1], fabs(e * zz)); ms[2] = fmax(ms[2], fabs(e * aa));
          vd[i] = okf * yd[i];
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) {                 // LOS rows
          double ax = 0.0;
          if (i < 3) {
#pragma unroll
            for (int j = 0; j < 4; ++j) ax = fma(K.C[4 * i + j], xk[r][j], ax);
          } else if (i == 3) {
            ax = c1 * xk[r][2] + c2 * xk[r][3];
          } else {
            ax = xk[r][1];
          }
          if (r < 2) ax = fma(K.Vecr[i], sk[r][i], ax);
          const double zi = i < 4 ? zs4[i] : zs5[0];
          const double e = Ev[4 * WRS_LOS(r, i)], pv = okf * (ax - zi), zz = okf * zi, aa = okf * ax;
          mu[0] = fmax(mu[0], fabs(pv)); mu[1] = fmax(mu[1], fabs(zz)); mu[2] = fmax(mu[2], fabs(aa));
          ms[0] = fmax(ms[0], fabs(e * pv)); ms[1] = fmax(ms[1], fabs(e * zz)); ms[2] = fmax(ms[2], fabs(e * aa));
          vs[i] = i < 4 ? okf * ys4[i] : 0.0;
        }
      } else {
        double dyd[4], dys[4] = {0.0, 0.0, 0.0, 0.0};
        wtm_ld4(tb + WTM_DD(r), dyd);
        if (r < 2) wtm_ld4(tb + WTM_DS(r), dys);
        const double bd0 = (r == 0 && c == 0) ? 1.0 : 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          vd[i] = dyd[i];
          vs[i] = dys[i];
          mu[6] = fmax(mu[6], fmax(fabs(dyd[i]), fabs(dys[i])));
          const double bnd = -prm[i] * bd0;                                   // dynamics rows: l = u
          lhs += bnd * dyd[i];
          lhs += los_hi(r, i) * fmax(dys[i], 0.0) + los_lo(r, i) * fmin(dys[i], 0.0);
        }
        vs[4] = 0.0;
      }
      at_stage(r, vd, vs);
      if (r < 2) {
        double vu[2], vbs[5];
        const double uf = st_u[r] ? 1.0 : 0.0, sf = st_s[r] ? 1.0 : 0.0;
        if (pass == 0) {
          double zu[2], yu[2], zb[4], yb[4], zyb[2];
          wtm_ld2(tb + WTM_ZU(r), zu);
          wtm_ld2(tb + WTM_YU(r), yu);
          wtm_ld4(tb + WTM_ZB(r), zb);
          wtm_ld2(tb + WTM_ZYB4(r), zyb);
          wtm_ld4(tb + WTM_YB(r), yb);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const double e = Ev[4 * WRS_BU(r, i)], pv = uf * (uk[r][i] - zu[i]), zz = uf * zu[i], aa = uf * uk[r][i];
            mu[0] = fmax(mu[0], fabs(pv)); mu[1] = fmax(mu[1], fabs(zz)); mu[2] = fmax(mu[2], fabs(aa));
            ms[0] = fmax(ms[0], fabs(e * pv)); ms[1] = fmax(ms[1], fabs(e * zz)); ms[2] = fmax(ms[2], fabs(e * aa));
            vu[i] = uf * yu[i];
          }
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const double zi = i < 4 ? zb[i] : zyb[0], yi = i < 4 ? yb[i] : zyb[1];
            const double e = Ev[4 * WRS_BS(r, i)], pv = sf * (sk[r][i] - zi), zz = sf * zi, aa = sf * sk[r][i];
            mu[0] = fmax(mu[0], fabs(pv)); mu[1] = fmax(mu[1], fabs(zz)); mu[2] = fmax(mu[2], fabs(aa));
            ms[0] = fmax(ms[0], fabs(e * pv)); ms[1] = fmax(ms[1], fabs(e * zz)); ms[2] = fmax(ms[2], fabs(e * aa));
            vbs[i] = sf * yi;
          }
        } else {
          double dyu[2], dyb[4], dyb4[2];
          wtm_ld2(tb + WTM_DU(r), dyu);
          wtm_ld4(tb + WTM_DB(r), dyb);
          wtm_ld2(tb + WTM_DB4(r), dyb4);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            vu[i] = dyu[i];
            mu[6] = fmax(mu[6], fabs(dyu[i]));
            lhs += K.ulim[i] * fmax(dyu[i], 0.0) - K.ulim[i] * fmin(dyu[i], 0.0);
          }
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            vbs[i] = i < 4 ? dyb[i] : dyb4[0];
            mu[6] = fmax(mu[6], fabs(vbs[i]));                                   // lower bound 0, no upper bound: no lhs term
          }
        }
        at_box(r, vd, vs, vu, vbs);
      }
    }
    {
      double pin4[4], dyp[2];
      if (pass == 0) {
        wtm_ld4(tb + WTM_PIN, pin4);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const double e = Ev[4 * WRS_PIN(i)], pv = dd[i] - pin4[i];
          mu[0] = fmax(mu[0], fabs(pv)); mu[1] = fmax(mu[1], fabs(pin4[i])); mu[2] = fmax(mu[2], fabs(dd[i]));
          ms[0] = fmax(ms[0], fabs(e * pv)); ms[1] = fmax(ms[1], fabs(e * pin4[i])); ms[2] = fmax(ms[2], fabs(e * dd[i]));
          vpin[i] = pin4[2 + i];
        }
      } else {
        wtm_ld2(tb + WTM_DP, dyp);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          vpin[i] = dyp[i];
          mu[6] = fmax(mu[6], fabs(dyp[i]));
          if (c == 0) lhs += prm[5 + i] * dyp[i];                                // carried by all four quad members: count once
        }
      }
    }
    // finish A' without the store of r: neighbour dynamics duals, d
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int kn = 4 * r + c + 1;
      const bool okn = kn <= NX;
      const int off = (c == 3) ? 16 * (r + 1) : 16 * r + (c + 1);
      double vn[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) vn[i] = okn ? vb[off + 4 * i] : 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double acc = rx[r][j];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double mij = (r == 0) ? K.Ad[4 * i + j] : (r == 2 ? K.Acl[4 * i + j] : M1nxt[4 * i + j]);
          acc = fma(mij, vn[i], acc);
        }
        rx[r][j] = acc;
      }
    }
    rd[0] += wq_sum(dsum[0]) + vpin[0];
    rd[1] += wq_sum(dsum[1]) + vpin[1];
    __syncwarp();
    if (pass == 0) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) aty_x[r][j] = rx[r][j];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int j = 0; j < 2; ++j) aty_u[r][j] = ru[r][j];
#pragma unroll
        for (int j = 0; j < 5; ++j) aty_s[r][j] = rs[r][j];
      }
      aty_d[0] = rd[0];
      aty_d[1] = rd[1];
    } else {                             // ||A' delta_y||_inf over the variables this thread owns
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (st_ok[r]) mu[7] = fmax(mu[7], fabs(rx[r][j]));
#pragma unroll
      for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (st_u[r]) mu[7] = fmax(mu[7], fabs(ru[r][j]));
#pragma unroll
        for (int j = 0; j < 5; ++j)
          if (st_s[r]) mu[7] = fmax(mu[7], fabs(rs[r][j]));
      }
      mu[7] = fmax(mu[7], fmax(fabs(rd[0]), fabs(rd[1])));
    }
  }
  // dual residual: c q + c P x + A' y per owned variable (unscaled; D-weighted for the adaptive-rho estimate)
  {
    auto dual = [&](double px_, double q_, double aty_, double dvv) {
      const double cpx = K.c * px_, dvl = q_ + cpx + aty_;
      mu[3] = fmax(mu[3], fabs(dvl)); mu[4] = fmax(mu[4], fabs(cpx)); mu[5] = fmax(mu[5], fabs(aty_));
      ms[3] = fmax(ms[3], fabs(dvv * dvl)); ms[4] = fmax(ms[4], fabs(dvv * cpx)); ms[5] = fmax(ms[5], fabs(dvv * aty_));
    };
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      if (st_ok[r]) {
        const bool fin = (r == 2 && c == 2);          // stage Nx carries the Riccati terminal weight
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          double px_ = 0.0;
#pragma unroll
          for (int j = 0; j < 4; ++j) px_ = fma(fin ? K.QN[4 * i + j] : K.Q[4 * i + j], xk[r][j], px_);
          dual(px_, fin ? K.qN[i] : K.qx[i], aty_x[r][i], Dv[4 * WVS_X(r, i)]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (st_u[r]) {
#pragma unroll
        for (int i = 0; i < 2; ++i) dual(K.Ru[2 * i] * uk[r][0] + K.Ru[2 * i + 1] * uk[r][1], 0.0, aty_u[r][i], Dv[4 * WVS_U(r, i)]);
      }
      if (st_s[r]) {
#pragma unroll
        for (int i = 0; i < 5; ++i) dual(K.Rs[i] * sk[r][i], 0.0, aty_s[r][i], Dv[4 * WVS_S(r, i)]);
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) dual(dd[i], 0.0, aty_d[i], Dv[4 * WVS_D(i)]);
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) mu[q] = wq_max(mu[q]);
  lhs = wq_sum(lhs);
  const double pri_u = mu[0], nz_u = mu[1], nax_u = mu[2], dua_u = mu[3] * K.cinv, npx_u = mu[4], naty_u = mu[5];
  const double ndy = mu[6], natdy = mu[7];
  auto check = [&](double kq) -> int {
    const double eps_p = kq * K.eps_abs + kq * K.eps_rel * fmax(nz_u, nax_u);
    const double eps_d = kq * K.eps_abs + kq * K.eps_rel * K.cinv * fmax(K.qn_unscaled, fmax(naty_u, npx_u));
    const bool prim_ok = pri_u < eps_p, dual_ok = dua_u < eps_d;
    if (prim_ok && dual_ok) return (kq > 1.0) ? 2 : 1;
    if (!prim_ok) {
      const double eps_i = kq * K.eps_pinf;
      if (ndy > MPCB_DIV_TOL && lhs < -eps_i * ndy && natdy < eps_i * ndy) return (kq > 1.0) ? 3 : -3;
    }
    return -10;
  };
#pragma unroll
  for (int q = 0; q < 6; ++q) ms[q] = wq_max(ms[q]);            // (shuffles stay outside the per-lane branches below)
  int st = check(1.0);
  if (st == -10) {
    if (K.adaptive && (iter % K.adapt_interval == 0)) {       // compute_rho_estimate on the SCALED residuals (OSQP 0.6.x)
      const double pr = ms[0] / (fmax(ms[1], ms[2]) + 1e-10);
      const double du = ms[3] / (fmax(K.qn_scaled, fmax(ms[5], ms[4])) + 1e-10);
      double est = rho * sqrt(pr / (du + 1e-10));
      est = fmin(fmax(est, MPCB_RHO_MIN), MPCB_RHO_MAX);
      if (est > rho * K.adapt_tol || est < rho / K.adapt_tol) rho = est;
    }
    if (iter >= K.max_iter) {
      st = check(10.0);
      if (st == -10) st = -2;
    }
  }

  // =============================== write back (OSQP-scaled format) ===============================
  // (the tensor-memory loads are warp-collective: every thread executes them, only the global stores are per lane)
  {
    double *xg = a.xs + (size_t)ln * N;
    double *zg = a.zs + (size_t)ln * M, *yg = a.ys + (size_t)ln * M;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int k = 4 * r + c;
      double zd[4], yd[4], zs4[4], ys4[4], zs5[2];
      wtm_ld4(tb + WTM_ZD(r), zd);
      wtm_ld4(tb + WTM_YD(r), yd);
      wtm_ld4(tb + WTM_ZS(r), zs4);
      wtm_ld2(tb + WTM_ZS4(r), zs5);
      wtm_ld4(tb + WTM_YS(r), ys4);
      if (valid && st_ok[r]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          xg[4 * k + j] = xk[r][j] * Dinv[4 * WVS_X(r, j)];
          zg[4 * k + j] = zd[j] * Ev[4 * WRS_DYN(r, j)];
          yg[4 * k + j] = yd[j] * Einv[4 * WRS_DYN(r, j)];
          zg[nX + 5 * k + j] = zs4[j] * Ev[4 * WRS_LOS(r, j)];
          yg[nX + 5 * k + j] = ys4[j] * Einv[4 * WRS_LOS(r, j)];
        }
        zg[nX + 5 * k + 4] = zs5[0] * Ev[4 * WRS_LOS(r, 4)];
        yg[nX + 5 * k + 4] = 0.0;
      }
      if (r < 2) {
        double zu[2], yu[2], zb[4], yb[4], zyb[2];
        wtm_ld2(tb + WTM_ZU(r), zu);
        wtm_ld2(tb + WTM_YU(r), yu);
        wtm_ld4(tb + WTM_ZB(r), zb);
        wtm_ld2(tb + WTM_ZYB4(r), zyb);
        wtm_ld4(tb + WTM_YB(r), yb);
        if (valid && st_u[r]) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            xg[nX + 7 * (k - 1) + j] = uk[r][j] * Dinv[4 * WVS_U(r, j)];
            zg[RB + 7 * (k - 1) + j] = zu[j] * Ev[4 * WRS_BU(r, j)];
            yg[RB + 7 * (k - 1) + j] = yu[j] * Einv[4 * WRS_BU(r, j)];
          }
        }
        if (valid && st_s[r]) {
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            xg[nX + 7 * k + 2 + j] = sk[r][j] * Dinv[4 * WVS_S(r, j)];
            zg[RB + 7 * k + 2 + j] = (j < 4 ? zb[j] : zyb[0]) * Ev[4 * WRS_BS(r, j)];
            yg[RB + 7 * k + 2 + j] = (j < 4 ? yb[j] : zyb[1]) * Einv[4 * WRS_BS(r, j)];
          }
        }
      }
    }
    double pin4[4];
    wtm_ld4(tb + WTM_PIN, pin4);
    if (valid && c == 3) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        xg[N - 2 + j] = dd[j] * Dinv[4 * WVS_D(j)];
        zg[M - 2 + j] = pin4[j] * Ev[4 * WRS_PIN(j)];
        yg[M - 2 + j] = pin4[2 + j] * Einv[4 * WRS_PIN(j)];
      }
    }
    if (valid && c == 0) {
      a.rho[ln] = rho;
      a.iter[ln] = iter;
      a.status[ln] = st;
      if (st != -10) a.lane_state[ln] = LANE_SOLVE_DONE;
    }
    if (valid && c == 1 && st != -10) {           // u_0 lives with stage 1: quad member 1, slot 0
      a.u0[ln] = uk[0][0];
      a.u0[Bz + ln] = uk[0][1];
    }
  }
  const unsigned vm = __ballot_sync(0xffffffffu, valid && c == 0);
  if (lid == 0 && vm) atomicAdd(a.iter_total, (unsigned long long)__popc(vm) * (unsigned long long)K.check_every);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(512));
}
