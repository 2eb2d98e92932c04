// Shared host/device structures and small device helpers for libmpcb (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MPCB_RHO_MIN 1e-6
#define MPCB_RHO_MAX 1e6
#define MPCB_RHO_EQ 1e3
#define MPCB_RHO_TOL 1e-4
#define MPCB_DIV_TOL 1e-30

// lane_state
enum : uint8_t { LANE_SOLVING = 0, LANE_SOLVE_DONE = 1, LANE_FINISHED = 2, LANE_DEFERRED = 3 };   // DEFERRED: left the rounds, waits for the team kernel

// Byte offsets of the sections of the per-variant constant blob staged into shared memory.
struct BlobHdr {
  int off_lam, off_q, off_D, off_Dinv, off_E, off_Einv, off_lt, off_ut;       // double vectors
  int off_Av, off_ATv, off_Pv;                                                // grouped-ELL values (double)
  int off_Ac, off_ATc, off_Pc;                                                // grouped-ELL columns (uint16)
  int off_Ag, off_ATg, off_Pg;                                                // int2 {offset,width} per 32-row group
  int off_flags;                                                              // uint8 per row: bit0 inf_l, bit1 inf_u, bit2 eq, bit3 free
  int off_V;                                                                  // V (n*n + pad doubles) or -1 if V stays in global
  int total;                                                                  // bytes (multiple of 16)
};

struct AdmmArgs {
  BlobHdr hdr;
  const unsigned char *blob[4];
  const double *Vg[4];
  int n, m, nX, Nx, Nb, uoff, B;
  double sigma, alpha, eps_abs, eps_rel, eps_pinf, adapt_tol, cinv, qn_unscaled, qn_scaled;
  int check_every, adaptive, adapt_interval, max_iter;
  const int *cnt;        // [4] lanes per variant this round
  const int *list;       // [4][B]
  double *xs, *zs, *ys;  // [B][n], [B][m], [B][m] lane-major ADMM iterates (scaled space)
  double *rho;           // [B]
  int *iter;             // [B] iterations done in the current solve
  int *status;           // [B] OSQP status_val of the current solve
  const double *par;     // [7][B] xhat0..3, val, dpin0, dpin1 (unscaled)
  double *u0;            // [2][B] first move of the last finished solve (unscaled)
  uint8_t *lane_state;   // [B]
  int *flip;             // [B] set when E*val < RHO_TOL on a row3 row (OSQP would reclassify it)
  unsigned long long *iter_total;  // [1] ADMM iterations executed (all lanes)
};

// Constant plant / estimator / controller data passed by value to the per-lane kernels.
struct SimConst {
  double Ad[16], Bd[8], Ao[36], Bou[12], Qw[36], Kpf[8], Kif[2], xr[4];
  double umax0, r_p, r_tol, suc_dist, suc_ang_deg, mean_mtn;
  int in_track, delta_v, is_reject, has_noise, noise_length;
  int estimator;               // MPCB_EST_UKF | MPCB_EST_KF
};

struct LaneSim {            // SoA [field][B] per-lane simulation state
  double *xtrue;            // [4][B]
  double *ux;               // [6][B]  UKF mean
  double *uP;               // [36][B] UKF covariance
  double *xstore;           // [4][B]  stored estimate (x/y swapped for in-track) used by the failsafe law
  double *uprev;            // [2][B]  ctrls[:, i]
  double *unext;            // [2][B]  ctrls[:, i+1]
  double *xintf;            // [B]
  double *noise;            // [2][B]
  double *xfin;             // [4][B]  x_true[:, i] of the last executed control step
  int *step;                // [B]
  int *sub;                 // [B] substep index (continuous simulator)
  int *iterm;               // [B]
  int *succ;                // [B]
  int *nsolve;              // [B]
  int *variant;             // [B]
  int *ukf_clamp;           // [B] set when a UKF Cholesky pivot was clamped to zero
};

struct SimOutDev {
  int32_t *i_term; int32_t *is_success; int32_t *ukf_clamped; double *final_dist;
  double *x_true, *x_est, *ctrl; uint8_t *ctrlr_seq; int8_t *status; int16_t *iters; double *u_raw;
  double *x_true_sub, *ctrl_sub; uint8_t *ctrlr_sub;   // continuous simulator, every substep: [.][NS][B]
  double *rho_hist;         // [T1-1][B] rho after solve i
  double *fd_all;           // [B] internal: every lane's final distance (NaN -> 0), summed in a fixed order by stats_fd_kernel
  int T1, NS;
};

enum : int { MODE_QP_ONLY = 0, MODE_DISCRETE = 1, MODE_CONTINUOUS = 2,
             MODE_RESUME = 3 };     // team kernel only: take over the listed lanes of a round-based discrete simulation mid-flight

struct PostArgs {
  SimConst sc;
  LaneSim ls;
  SimOutDev out;
  int mode, B, nsteps;
  // continuous simulator
  int ratio, n_sub_total, noise_hold_sub;
  double T_cont;
  const double *noise_in;   // [n_refresh][2][B]
  int n_refresh;
  // QP coupling
  double *par;              // [7][B]
  const double *u0;         // [2][B]
  double *rho;
  int *iter, *status;
  uint8_t *lane_state;
  int *cnt_cur, *cnt_next;  // [4] each
  int *list_next;           // [4][B]
  // lanes whose next solve has re-typed rows (|p^ - r|_1 < defer_below <=> some velocity-bound row's scaled bounds come within
  // RHO_TOL, team.cuh tm_retype_operator) leave the rounds: they are listed here and carried to the end of their
  // trajectories by the team kernel (MODE_RESUME), the one solver block that folds re-typed rows into its operator
  double defer_below;       // <= 0: no deferral
  int *cnt_def, *list_def;  // [4], [4][B]
  unsigned long long *solves_total;
};

// sum final_dist and final_dist^2 over the batch in an order that depends on B only (one block, strided partial sums, fixed
// tree): run-to-run identical, unlike atomicAdd(double) from the lanes as they finish.  The other statistics are integer
// counts, whose double sums are exact in any order.
__global__ void __launch_bounds__(1024) stats_fd_kernel(const double *__restrict__ fd, int B, double *__restrict__ stats) {
  __shared__ double s0[1024], s1[1024];
  double a0 = 0.0, a1 = 0.0;
  for (int i = threadIdx.x; i < B; i += 1024) {
    const double f = fd[i];
    a0 += f;
    a1 += f * f;
  }
  s0[threadIdx.x] = a0;
  s1[threadIdx.x] = a1;
  __syncthreads();
  for (int w = 512; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) {
      s0[threadIdx.x] += s0[threadIdx.x + w];
      s1[threadIdx.x] += s1[threadIdx.x + w];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    stats[0] = s0[0];
    stats[1] = s1[0];
  }
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// max over the warp of NON-NEGATIVE doubles with two REDUX instructions: for v >= 0 the IEEE bit pattern orders like
// the (hi, lo) pair of unsigned words, so reduce hi, then lo among the lanes that hold the winning hi.
__device__ __forceinline__ double warp_max_nonneg(double v) {
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  return __hiloint2double((int)mh, (int)ml);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- mbarrier + bulk async copy (TMA engine, non-tensor form), sm_90+/sm_100a PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
