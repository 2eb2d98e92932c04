// Team kernel: one CTA ("team", 256 or 384 threads) owns one trajectory at a time.  Discrete simulator: the team runs
// the trajectory's WHOLE closed loop (every control step: OSQP-equivalent ADMM solve -> controller select / clip ->
// plant step -> UKF -> QP parameter refresh) without returning to the host.  Round-based simulators (continuous
// plant): "list mode", the team solves the QP of every lane in this round's lists.  Teams pull lanes from an atomic
// queue until the batch / the lists are exhausted (persistent grid, 2 teams per SM at Nx = 10 / 20, 1 at Nx = 30).
//
// Linear solve.  OSQP refactors its KKT matrix whenever rho changes (reference
// src/trajectorySimulate.py:296 -> osqp_solve / adapt_rho).  Here the reduced operator
//     S(rho) = (P + sigma I + A' diag(rho_vec) A)^-1 = V diag(1/(1+rho*lam)) V'
// (spectral tables V, lam per sign variant, built once on the host, problem.py) is materialised per trajectory ON
// CHIP: threads 2i and 2i+1 own the two halves of row i of S (HALF doubles each), so one ADMM iteration is ONE dense
// mat-vec.  TM = true (default): S lives in TENSOR MEMORY, 2*HALF 32-bit columns of the thread's own TMEM lane
// (tcgen05.st at a rebuild, tcgen05.ld in 8-column chunks inside the mat-vec) and the registers hold the thread's
// rows of A / A'; TM = false: S in registers (+ SS entries per thread in shared memory), A / A' in shared memory.
// r is broadcast from shared memory with 128-bit loads and the two half sums meet in a shuffle.  S is rebuilt
// (N^3 FMAs across the team) only when the trajectory's rho adapts or its velocity-sign variant flips; in list mode
// it is parked in / reloaded from a per-lane cache in HBM between visits.
//
// Per iteration (3 team barriers):
//     r  = sigma*x - q + A'v        (thread pair per entry of r, ELL)      -> smem, barrier
//     xt = S r ; x = a*xt+(1-a)*x   (thread pair per row of S)             -> smem, barrier
//     zt = A xt ; z,y update ; v = rho_vec.*z - y  (thread per row, ELL)   -> smem, barrier
#pragma once
#include "common.cuh"
#include "sim.cuh"

#ifndef TEAM_OPC
#define TEAM_OPC 3             // operators kept per team in the L2 cache (TeamArgs::ocache)
#endif
#ifndef TEAM_TM_CHUNK
#define TEAM_TM_CHUNK 8        // TMEM columns per tcgen05.ld in the mat-vec (8 or 16)
#endif

// Optional cycle breakdown (compile with -DTEAM_PROFILE): thread 0 of every team accumulates clock64()
// deltas per phase into tot[4..]: 4 iterations, 5 checks, 6 operator rebuilds, 7 post step, 8 lane setup, 9 total.
#ifdef TEAM_PROFILE
#define PS_T0 long long ps_t = clock64();
#define UK_T0 long long uk_t = clock64();
#define UK_MARK(k) { const long long uk_n = clock64(); if (lid == 0) atomicAdd(&w.dbg[(k)], (unsigned long long)(uk_n - uk_t)); uk_t = uk_n; }
#define PS_MARK(k) { const long long ps_n = clock64(); if (lid == 0) atomicAdd(&a.tot[10 + (k)], (unsigned long long)(ps_n - ps_t)); ps_t = ps_n; }
#ifndef TP_TID
#define TP_TID 0
#endif
#define TP_DECL long long tp_t0 = clock64(), tp_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long tp_ph = 0; const long long tp_start = tp_t0;
#define TP_MARK(slot) { const long long tp_now = clock64(); tp_acc[slot] += tp_now - tp_t0; tp_t0 = tp_now; }
#define TP_PH0 tp_ph = clock64();
#define TP_PH(slot) { const long long tp_now = clock64(); tp_acc[slot] += tp_now - tp_ph; tp_ph = tp_now; }
#define TP_FLUSH if (tid == TP_TID) { for (int q = 5; q < 11; ++q) atomicAdd(&a.tot[q + 5], (unsigned long long)tp_acc[q]); } if (tid == 0) { for (int q = 0; q < 5; ++q) atomicAdd(&a.tot[4 + q], (unsigned long long)tp_acc[q]); \
                                 atomicAdd(&a.tot[9], (unsigned long long)(clock64() - tp_start)); }
#else
#define TP_DECL
#define TP_MARK(slot)
#define PS_T0
#define PS_MARK(k)
#define UK_T0
#define UK_MARK(k)
#define TP_PH0
#define TP_PH(slot)
#define TP_FLUSH
#endif

// Byte offsets into the team constant blob (global -> shared once per CTA).
//
// Sparse tables are ELL in "team layout".  Rows of A are handed to threads in order of decreasing
// nnz (thread t <-> row rowmap[t]), columns of A (= rows of A') to thread PAIRS likewise
// (pair p <-> variable colmap[p]; the pair splits the column's entries even/odd).  Every warp then
// has its own ELL width (wA[warp], wAT2[warp]) and skips the padding of the widest rows: shared-
// memory wavefronts, not FP64 issue, bound this kernel (profiles/team_kernel_r1.md).
struct TeamHdr {
  int off_q, off_D, off_Dinv, off_E, off_Einv, off_lt, off_ut;   // double vectors in natural index order
  int off_Av;       // [WA][MP]    value e of the row thread t owns
  int off_ATv;      // [WAT2+WATX][NCT] value pair-slot e of thread t (entry 2e + (t&1) of column colmap[t>>1])
  int off_Pv;       // [WP][NCT/2] row colmap[p] of P
  int off_Ac;       // [MP] uint4: 8 uint16 column indices of the thread's row
  int off_ATc;      // [NCT] uint4: 8 uint16 row indices (pair-slots 0..7)
  int off_ATc2;     // [NCT] uint4: pair-slots 8..15 (long columns: the disturbance variables touch every dynamics row)
  int off_Pc;       // [WP][NCT/2] uint16
  int off_rowmap;   // [MP] uint16
  int off_colmap;   // [NCT/2] uint16
  int off_flags;                                                 // uint8 per row (natural order)
  int off_patch;                                                 // int4 {offA, offAT, which, 0} per signed entry
  int n_patch;
  int wA[16], wAT2[16];                                          // per-warp ELL widths
  int r3c[8];                                                    // velocity-bound row of stage k: column of its third entry (the slack; -1: none)
  double r3v[8];                                                 //   and that entry's scaled value (sign-variant independent); tm_retype_operator
  int r3ok;                                                      // 0: rows do not have the expected shape -> re-typing is only counted
  int total;                                                     // bytes, multiple of 16
};

struct TeamArgs {
  TeamHdr hdr;
  const unsigned char *blob;
  const double *Vk[4];       // [N][NP2] k-major: Vk[k*NP2 + j] = V[j][k], zero padded
  const double *lam;         // [4][N]
  int n, m, nX, Nb, uoff, B;
  double sigma, alpha, eps_abs, eps_rel, eps_pinf, adapt_tol, cinv, qn_unscaled, qn_scaled, rho0;
  int check_every, adaptive, adapt_interval, max_iter;
  int mode;                  // MODE_QP_ONLY | MODE_DISCRETE | MODE_RESUME
  int nsteps;
  LaneSim ls;                // MODE_RESUME: the round-based simulators' per-lane state (sim.cuh), read at lane start and written back at
                             // the lane's end; the listed lanes (cnt / list) are carried to the end of their trajectories
  SimConst sc;
  SimOutDev out;
  const double *x0;          // [4][B]
  const double *noise_in;    // [n_refresh][2][B]
  int n_refresh;
  // QP seam / persistent solver state
  const double *xhat;        // [6][B] (MODE_QP_ONLY)
  double *xs, *zs, *ys;      // [B][n], [B][m], [B][m]
  double *rho;               // [B]
  double *u0;                // [2][B]
  int *iter, *status, *flip;
  int warm;                  // 1: load x,z,y,rho from global (QP seam), 0: cold start
  // list mode (MODE_QP_ONLY with list != nullptr): the round-based simulators' contract (admm.cuh) -- solve every lane
  // of this round's per-variant lists to completion; parameters come from par, the verdict goes to lane_state
  const int *cnt;            // [4] lanes per sign variant
  const int *list;           // [4][B]
  const double *par;         // [7][B]
  uint8_t *lane_state;       // [B]
  int visit_iters;           // > 0: at most this many iterations per visit; an unfinished lane keeps status -10 and stays listed,
                             //      so one 4000-iteration solve does not hold up a round of 25..100-iteration ones
  // per-lane operator cache (tensor-memory kernels): S of the lane's last solve, tagged with (rho, variant), so that a lane
  // revisited every control step reloads 64 KB from HBM instead of redoing the n^3 rebuild
  double *scache;            // [B][SCACHE_CHUNKS][NCT] double2, or nullptr
  double *scache_rho;        // [B] rho the cached S was built for (< 0: empty)
  int *scache_var;           // [B]
  // per-TEAM operator cache (whole-loop modes): the last TEAM_OPC operators this team built, [grid][TEAM_OPC][chunks][NCT]
  // double2 in global memory (L2-resident).  Trajectories flip between velocity-sign variants at unchanged rho (95 % of the
  // rebuilds of config 4); a flip back costs a 64-231 KB reload instead of the n^3 rebuild.
  double *ocache;
  // results
  int *queue;                // [1] next lane
  unsigned long long *tot;   // [0] admm iterations, [1] qp solves, [2] operator rebuilds
  double *stats;             // [MPCB_NSTATS]
};

struct LaneCtx {
  double xtrue[4], ux[6], uP[36], xstore[4], unext[2], noise[2], xfin[4], par[7], u0[2];
  double xintf, rho;
  int step, iterm, succ, nsolve, nsolve0, variant, ukf_clamp, flip, fin, status, iter, lane;
  double c_rho;              // operator-cache tag of this lane
  int c_var;
};

struct UkfScratch {
  double Ao[36], Bou[12], Qw[36];      // shared copies: per-thread indexed reads of kernel parameters serialise
  double U[36], sig[78], sf[78], xm[6], Pm[36], zs[26], zp[2], S[4], Pxz[12], K[12];
  int ok;
  unsigned long long *dbg;             // TEAM_PROFILE builds: cycle counters
};

__device__ __forceinline__ int ctx_params(const SimConst &c, LaneCtx &L, const double *xe) {
  L.par[0] = xe[0]; L.par[1] = xe[1]; L.par[2] = xe[2]; L.par[3] = xe[3];
  L.par[4] = fabs(xe[0] - c.xr[0]) + fabs(xe[1] - c.xr[1]);
  L.par[5] = c.is_reject ? xe[4] : 0.0;
  L.par[6] = c.is_reject ? xe[5] : 0.0;
  return (xe[2] >= 0 ? 0 : 1) + (xe[3] >= 0 ? 0 : 2);          // sign(0) = +1, simhelpers.py:66-67
}

// kf.predict(u); kf.update(z) (filterpy 1.4.5 as restated in oracle/ukf_ref.py, R = 0) executed by
// ONE WARP on shared-memory operands; same operation order per output as the scalar ukf_step().
__device__ __forceinline__ void ukf_sigma_warp(const double *x, const double *U, double *sig, int lid) {
  for (int idx = lid; idx < 78; idx += 32) {
    const int k = idx / 6, j = idx - 6 * k;
    sig[idx] = (k == 0) ? x[j] : ((k <= 6) ? x[j] + U[(k - 1) * 6 + j] : x[j] - U[(k - 7) * 6 + j]);
  }
}

// chol_upper6 with the matrix pulled into registers first and every loop unrolled: the shared-memory
// version serialises ~100 dependent LDS round trips, and sqrt + divisions cost ~5 000 cycles per call.
__device__ __forceinline__ bool chol_upper6_reg(const double *Psm, double s, double *Usm) {
  double P[36], U[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) {
    P[i] = Psm[i];
    U[i] = 0.0;
  }
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double d = s * P[i * 6 + i];
#pragma unroll
    for (int k = 0; k < i; ++k) d -= U[k * 6 + i] * U[k * 6 + i];
    if (d > 0.0) {
      const double rinv = rsqrt(d);             // one reciprocal square root instead of a sqrt and 5 divisions:
      U[i * 6 + i] = d * rinv;                  // differs from sqrt(d), v / sqrt(d) by an ulp
#pragma unroll
      for (int j = i + 1; j < 6; ++j) {
        double v = s * P[i * 6 + j];
#pragma unroll
        for (int k = 0; k < i; ++k) v -= U[k * 6 + i] * U[k * 6 + j];
        U[i * 6 + j] = v * rinv;
      }
    } else {
      ok = false;                               // row i stays zero (see chol_upper6)
    }
  }
#pragma unroll
  for (int i = 0; i < 36; ++i) Usm[i] = U[i];
  return ok;
}

__device__ __noinline__ bool ukf_step_warp(const SimConst &c, double *x, double *P, double u0, double u1, double z0, double z1,
                                           UkfScratch &w, int lid) {
  // predict.  The process model is linear (fx = Ao x + Bou u, trajectorySimulate.py:121-123), so its
  // unscented transform is evaluated directly: x- = Ao x + Bou u, P- = Ao P Ao' + Q.  This is what
  // the 13 propagated sigma points sum to (the +/- pairs cancel every cross term; the centre point
  // contributes Wc0 * eps eps' with eps ~ 1e-12) up to rounding at the 1e-16 level, without the
  // first Cholesky, the 78 propagations and the two 13-term weighted sums.
  UK_T0
  if (lid == 0) w.ok = 1;
  if (lid < 6) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) acc += w.Ao[lid * 6 + j] * x[j];
    w.xm[lid] = acc + w.Bou[lid * 2] * u0 + w.Bou[lid * 2 + 1] * u1;
  }
  for (int idx = lid; idx < 36; idx += 32) {          // T = Ao P   (kept in w.sig)
    const int i = idx / 6, j = idx - 6 * i;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) acc += w.Ao[i * 6 + k] * P[k * 6 + j];
    w.sig[idx] = acc;
  }
  __syncwarp();
  for (int idx = lid; idx < 36; idx += 32) {          // P- = T Ao' + Q
    const int i = idx / 6, j = idx - 6 * i;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) acc += w.sig[i * 6 + k] * w.Ao[j * 6 + k];
    w.Pm[idx] = acc + w.Qw[idx];
  }
  __syncwarp();
  UK_MARK(13)
  // Two Cholesky factorisations side by side: lane 0 factors 0.05 P- (the sigma points of the update),
  // lane 1 factors 0.05 P only to learn whether filterpy's predict would have raised on a non-positive
  // pivot (the lane is reported).  Same instruction stream, different data.
  if (lid < 2) {
    const bool ok = chol_upper6_reg(lid == 0 ? w.Pm : P, UKF_NPL, lid == 0 ? w.U : w.sf);
    if (!ok) w.ok = 0;
  }
  __syncwarp();
  UK_MARK(14)
  ukf_sigma_warp(w.xm, w.U, w.sf, lid);          // filterpy 1.4.5 regenerates the points after predict
  __syncwarp();
  if (lid < 13) {
    const double a = w.sf[lid * 6], b = w.sf[lid * 6 + 1];
    w.zs[lid * 2] = sqrt(a * a + b * b);
    w.zs[lid * 2 + 1] = atan2(b, a);
  }
  __syncwarp();
  UK_MARK(15)
  if (lid < 2) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 13; ++k) acc += (k == 0 ? UKF_WM0 : UKF_WI) * w.zs[k * 2 + lid];
    w.zp[lid] = acc;
  }
  __syncwarp();
  if (lid < 16) {
    double acc = 0.0;
    if (lid < 4) {
      const int r = lid >> 1, q = lid & 1;
#pragma unroll
      for (int k = 0; k < 13; ++k)
        acc += (k == 0 ? UKF_WC0 : UKF_WI) * (w.zs[k * 2 + r] - w.zp[r]) * (w.zs[k * 2 + q] - w.zp[q]);
      w.S[lid] = acc;
    } else {
      const int e = lid - 4, i = e >> 1, q = e & 1;
#pragma unroll
      for (int k = 0; k < 13; ++k)
        acc += (k == 0 ? UKF_WC0 : UKF_WI) * (w.sf[k * 6 + i] - w.xm[i]) * (w.zs[k * 2 + q] - w.zp[q]);
      w.Pxz[e] = acc;
    }
  }
  __syncwarp();
  const double det = w.S[0] * w.S[3] - w.S[1] * w.S[2];
  const double SI[4] = {w.S[3] / det, -w.S[1] / det, -w.S[2] / det, w.S[0] / det};
  if (lid < 12) {
    const int i = lid >> 1, q = lid & 1;
    w.K[lid] = w.Pxz[i * 2] * SI[q] + w.Pxz[i * 2 + 1] * SI[2 + q];
  }
  __syncwarp();
  const double y0 = z0 - w.zp[0], y1 = z1 - w.zp[1];
  if (lid < 6) x[lid] = w.xm[lid] + w.K[lid * 2] * y0 + w.K[lid * 2 + 1] * y1;
  for (int idx = lid; idx < 36; idx += 32) {
    const int i = idx / 6, j = idx - 6 * i;
    const double ks0 = w.K[i * 2] * w.S[0] + w.K[i * 2 + 1] * w.S[2], ks1 = w.K[i * 2] * w.S[1] + w.K[i * 2 + 1] * w.S[3];
    P[idx] = w.Pm[idx] - (ks0 * w.K[j * 2] + ks1 * w.K[j * 2 + 1]);
  }
  __syncwarp();
  return w.ok != 0;
}

// The rest of one control step after the solve (trajectorySimulate.py:298-356), executed by warp 0
// of the team on the lane context in shared memory (scalar parts by its lane 0).
__device__ __noinline__ void lane_post_step(const TeamArgs &a, LaneCtx &L, UkfScratch &w, int lid) {
  const SimConst &c = a.sc;
  const size_t B = a.B;
  const int ln = L.lane, T1 = a.out.T1, i = L.step;
  const double up0 = L.unext[0], up1 = L.unext[1];               // ctrls[:, i], the command chosen one step ago
  double xn[4];
  PS_T0
  plant_lin(c, L.xtrue, L.unext, L.noise, xn);                   // every lane of the warp: cheap, avoids a broadcast
  __syncwarp();
  if (lid == 0) {
    double u[2], uraw[2];
    int code;
    if (L.status != 1) {                                         // failsafe LQR with integrator (:305-309)
      const double xi = L.xintf + L.xstore[0] - c.xr[0];
      L.xintf = xi;
      for (int r = 0; r < 2; ++r) {
        double acc = 0.0;
        for (int j = 0; j < 4; ++j) acc += c.Kpf[r * 4 + j] * L.xstore[j];
        u[r] = -acc - c.Kif[r] * xi;
      }
      code = 2;
    } else {
      L.xintf = 0.0;
      u[0] = L.u0[0];
      u[1] = L.u0[1];
      code = 1;
    }
    uraw[0] = u[0];
    uraw[1] = u[1];
    const double nrm = sqrt(u[0] * u[0] + u[1] * u[1]);
    if (nrm > c.umax0) {                                         // sequential clip (:317-319)
      u[0] = u[0] * (c.umax0 / nrm);
      const double nrm2 = sqrt(u[0] * u[0] + u[1] * u[1]);
      u[1] = u[1] * (c.umax0 / nrm2);
    }
    L.nsolve += 1;
    if (a.out.status) a.out.status[(size_t)i * B + ln] = (int8_t)L.status;
    if (a.out.iters) a.out.iters[(size_t)i * B + ln] = (int16_t)L.iter;
    if (a.out.rho_hist) a.out.rho_hist[(size_t)i * B + ln] = L.rho;
    if (a.out.ctrlr_seq) a.out.ctrlr_seq[(size_t)i * B + ln] = (uint8_t)code;
    if (a.out.u_raw) {
      a.out.u_raw[((size_t)0 * (T1 - 1) + i) * B + ln] = uraw[0];
      a.out.u_raw[((size_t)1 * (T1 - 1) + i) * B + ln] = uraw[1];
    }
    if (a.out.ctrl) {
      a.out.ctrl[((size_t)0 * T1 + i + 1) * B + ln] = u[0];
      a.out.ctrl[((size_t)1 * T1 + i + 1) * B + ln] = u[1];
    }
    L.unext[0] = u[0];
    L.unext[1] = u[1];
    if (i >= 1 && success_cond(c, L.xtrue)) L.succ = 1;          // scan over x_true[:, 1 .. i_term-1] (:369-376)
    for (int k = 0; k < 4; ++k) L.xfin[k] = L.xtrue[k];
  }
  __syncwarp();
  PS_MARK(0)
  double xe[6];
  if (c.has_noise) {
    if (c.estimator == MPCB_EST_KF) {          // linear KF: a few hundred flops, one lane does it
      if (lid == 0) {
        const double uu[2] = {up0, up1}, yy[2] = {xn[0], xn[1]};
        kf_step(c, L.ux, L.uP, uu, yy);
      }
      __syncwarp();
    } else {
      const double z0 = sqrt(xn[0] * xn[0] + xn[1] * xn[1]), z1 = atan2(xn[1], xn[0]);
      const bool ok = ukf_step_warp(c, L.ux, L.uP, up0, up1, z0, z1, w, lid);
      if (!ok && lid == 0) L.ukf_clamp = 1;
    }
    for (int k = 0; k < 6; ++k) xe[k] = L.ux[k];
  } else {
    for (int k = 0; k < 4; ++k) xe[k] = xn[k];
    xe[4] = xe[5] = 0.0;
  }
  __syncwarp();
  PS_MARK(1)
  if (lid == 0) {
    L.variant = ctx_params(c, L, xe);
    if (c.in_track) {                                            // in-place x/y swap of the stored estimate, simhelpers.py:72
      const double t = xe[0];
      xe[0] = xe[1];
      xe[1] = t;
    }
    for (int k = 0; k < 4; ++k) L.xstore[k] = xe[k];
    if (a.out.x_est)
      for (int k = 0; k < 6; ++k) a.out.x_est[((size_t)k * T1 + i + 1) * B + ln] = xe[k];
    if (a.out.x_true)
      for (int k = 0; k < 4; ++k) a.out.x_true[((size_t)k * T1 + i + 1) * B + ln] = xn[k];
    for (int k = 0; k < 4; ++k) L.xtrue[k] = xn[k];
    if (c.has_noise && ((i + 1) % c.noise_length == 0)) {
      const int r = min((i + 1) / c.noise_length, a.n_refresh - 1);
      L.noise[0] = a.noise_in[((size_t)r * 2 + 0) * B + ln];
      L.noise[1] = a.noise_in[((size_t)r * 2 + 1) * B + ln];
    }
    L.step = i + 1;
    if (i + 1 >= a.nsteps) {
      L.fin = 1;                                                 // i_term stays nsim
    } else if (terminated(c, xn)) {
      L.iterm = i + 1;
      L.fin = 1;
    }
  }
  __syncwarp();
  PS_MARK(2)
}

__device__ __noinline__ void lane_init(const TeamArgs &a, LaneCtx &L, int ln) {
  const SimConst &c = a.sc;
  const size_t B = a.B;
  L.lane = ln;
  L.rho = a.rho0;
  L.xintf = 0.0;
  L.step = 0; L.succ = 0; L.nsolve = 0; L.ukf_clamp = 0; L.flip = 0; L.fin = 0; L.status = -10; L.iter = 0;
  L.u0[0] = L.u0[1] = 0.0;
  L.c_rho = -1.0;
  L.c_var = -1;
  L.nsolve0 = 0;
  if (a.mode == MODE_RESUME) {           // mid-flight lane of a round-based run (variant came with the list it was drawn from)
    for (int k = 0; k < 7; ++k) L.par[k] = a.par[k * B + ln];
    for (int k = 0; k < 4; ++k) {
      L.xtrue[k] = a.ls.xtrue[k * B + ln];
      L.xstore[k] = a.ls.xstore[k * B + ln];
      L.xfin[k] = a.ls.xfin[k * B + ln];
    }
    for (int k = 0; k < 6; ++k) L.ux[k] = a.ls.ux[k * B + ln];
    for (int k = 0; k < 36; ++k) L.uP[k] = a.ls.uP[k * B + ln];
    for (int k = 0; k < 2; ++k) {
      L.unext[k] = a.ls.unext[k * B + ln];
      L.noise[k] = a.ls.noise[k * B + ln];
    }
    L.xintf = a.ls.xintf[ln];
    L.rho = a.rho[ln];
    L.iter = a.iter[ln];                 // iterations the lane's current solve has already done
    L.step = a.ls.step[ln];
    L.iterm = a.ls.iterm[ln];
    L.succ = a.ls.succ[ln];
    L.nsolve = L.nsolve0 = a.ls.nsolve[ln];
    L.ukf_clamp = a.ls.ukf_clamp[ln];
    L.flip = a.flip[ln];
    return;
  }
  if (a.mode == MODE_QP_ONLY) {
    if (a.list) {            // variant was set by the caller (it comes with the list the lane was drawn from)
      for (int k = 0; k < 7; ++k) L.par[k] = a.par[k * B + ln];
      L.rho = a.rho[ln];
      L.iter = a.iter[ln];
      if (a.scache) { L.c_rho = a.scache_rho[ln]; L.c_var = a.scache_var[ln]; }
    } else {
      double xe[6];
      for (int k = 0; k < 6; ++k) xe[k] = a.xhat[k * B + ln];
      L.variant = ctx_params(c, L, xe);
      if (a.warm) L.rho = a.rho[ln];
    }
    L.iterm = 0;
    return;
  }
  double x[6];
  for (int k = 0; k < 4; ++k) x[k] = a.x0[k * B + ln];
  x[4] = x[5] = 0.0;
  for (int k = 0; k < 4; ++k) {
    L.xtrue[k] = x[k];
    L.xstore[k] = x[k];
    L.xfin[k] = nan("");
  }
  for (int k = 0; k < 6; ++k) L.ux[k] = x[k];
  for (int k = 0; k < 36; ++k) L.uP[k] = (k % 7 == 0) ? ((k / 7 < 4) ? 1e-20 : 1.0) : 0.0;
  L.unext[0] = L.unext[1] = 0.0;
  L.noise[0] = (c.has_noise && a.noise_in) ? a.noise_in[ln] : 0.0;
  L.noise[1] = (c.has_noise && a.noise_in) ? a.noise_in[B + ln] : 0.0;
  L.variant = ctx_params(c, L, x);
  const int T1 = a.out.T1;
  if (a.out.x_true) for (int k = 0; k < 4; ++k) a.out.x_true[((size_t)k * T1) * B + ln] = x[k];
  if (a.out.x_est) for (int k = 0; k < 6; ++k) a.out.x_est[((size_t)k * T1) * B + ln] = x[k];
  if (a.out.ctrl) for (int k = 0; k < 2; ++k) a.out.ctrl[((size_t)k * T1) * B + ln] = 0.0;
  L.iterm = a.nsteps;
  if (a.nsteps <= 0) L.fin = 1;
  else if (terminated(c, x)) { L.iterm = 0; L.fin = 1; }
}

__device__ __noinline__ void lane_finalize(const TeamArgs &a, LaneCtx &L) {
  const int ln = L.lane;
  if (a.mode == MODE_RESUME) {           // hand the finished lane back: finalize_kernel reduces every lane from LaneSim
    const size_t B = a.B;
    for (int k = 0; k < 4; ++k) {
      a.ls.xfin[k * B + ln] = L.xfin[k];
      a.ls.xtrue[k * B + ln] = L.xtrue[k];
    }
    a.ls.iterm[ln] = L.iterm;
    a.ls.succ[ln] = L.succ;
    a.ls.nsolve[ln] = L.nsolve;
    a.ls.step[ln] = L.step;
    a.ls.ukf_clamp[ln] = L.ukf_clamp;
    a.rho[ln] = L.rho;
    if (L.flip) a.flip[ln] = 1;
    a.lane_state[ln] = LANE_FINISHED;
    atomicAdd(&a.tot[1], (unsigned long long)(L.nsolve - L.nsolve0));
    return;
  }
  if (a.mode == MODE_QP_ONLY) {
    a.status[ln] = L.status;
    a.iter[ln] = L.iter;
    a.u0[ln] = L.u0[0];
    a.u0[(size_t)a.B + ln] = L.u0[1];
    a.rho[ln] = L.rho;
    if (L.flip) a.flip[ln] = 1;
    if (a.list) {
      if (L.status != -10) a.lane_state[ln] = LANE_SOLVE_DONE;
      if (a.scache) { a.scache_rho[ln] = L.c_rho; a.scache_var[ln] = L.c_var; }
    } else {
      atomicAdd(&a.tot[1], 1ull);
    }
    return;
  }
  double d2 = 0.0;
  for (int k = 0; k < 4; ++k) {
    const double d = L.xfin[k] - a.sc.xr[k];
    d2 += d * d;
  }
  const double fd = sqrt(d2);
  if (a.out.i_term) a.out.i_term[ln] = L.iterm;
  if (a.out.is_success) a.out.is_success[ln] = L.succ;
  if (a.out.final_dist) a.out.final_dist[ln] = fd;
  if (a.out.ukf_clamped) a.out.ukf_clamped[ln] = L.ukf_clamp;
  a.rho[ln] = L.rho;
  const double f = (fd == fd) ? fd : 0.0;
  if (a.out.fd_all) a.out.fd_all[ln] = f;         // stats[0], stats[1]: summed in a fixed order afterwards (stats_fd_kernel)
  if (L.succ) atomicAdd(&a.stats[2], 1.0);
  atomicAdd(&a.stats[3], 1.0);
  atomicAdd(&a.stats[4], (double)L.iterm);
  atomicAdd(&a.stats[5], (double)L.nsolve);
  if (L.flip) atomicAdd(&a.stats[7], 1.0);
  if (L.ukf_clamp) atomicAdd(&a.stats[8], 1.0);
  if (L.iterm < a.nsteps) atomicAdd(&a.stats[9], 1.0);
  atomicAdd(&a.tot[1], (unsigned long long)L.nsolve);
}

// W entries of one ELL row: values at vals[e*stride], 16-bit column indices packed in c.
template <int W>
__device__ __forceinline__ double ell_dot(const double *vals, int stride, const uint4 &c, const double *vec) {
  const unsigned cw[4] = {c.x, c.y, c.z, c.w};
  double v[W], g[W];
#pragma unroll
  for (int e = 0; e < W; ++e) {
    v[e] = vals[e * stride];
    g[e] = vec[(e & 1) ? (cw[e >> 1] >> 16) : (cw[e >> 1] & 0xffffu)];
  }
  double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
  for (int e = 0; e < W; ++e) {
    if (e & 1) acc1 = fma(v[e], g[e], acc1);
    else acc0 = fma(v[e], g[e], acc0);
  }
  return acc0 + acc1;
}


// W entries of one ELL row whose values / packed column indices already sit in registers (tensor-memory mode).
template <int W, int WMAX>
__device__ __forceinline__ double ell_dot_reg(const double (&v)[WMAX], const uint4 &c, const double *vec) {
  const unsigned cw[4] = {c.x, c.y, c.z, c.w};
  double g[W];
#pragma unroll
  for (int e = 0; e < W; ++e) g[e] = vec[(e & 1) ? (cw[e >> 1] >> 16) : (cw[e >> 1] & 0xffffu)];
  double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
  for (int e = 0; e < W; ++e) {
    if (e & 1) acc1 = fma(v[e], g[e], acc1);
    else acc0 = fma(v[e], g[e], acc0);
  }
  return acc0 + acc1;
}

// ---- tensor memory as per-thread operand storage: shape 32x32b, thread t of a warp <-> TMEM lane 32*(warp%4)+t,
//      consecutive 32-bit columns <-> consecutive registers.  The per-lane operator S lives here in TM mode.
#define TMEM_LD_REGS16(r) "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
#define TMEM_ST_REGS16(r) "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), \
                          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : TMEM_LD_REGS16(r) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld8x2(uint32_t (&r)[8], uint32_t (&q)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(q[0]), "+r"(q[1]),
                 "+r"(q[2]), "+r"(q[3]), "+r"(q[4]), "+r"(q[5]), "+r"(q[6]), "+r"(q[7]) :: "memory");
}
__device__ __forceinline__ void tmem_wait_ld8(uint32_t (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) :: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), TMEM_ST_REGS16(r) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
// The waits name the registers the pending loads write ("+r"): the compiler must not move a read of them above the wait.
#define TMEM_RW_REGS16(r) "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), \
                          "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
__device__ __forceinline__ void tmem_wait_ld2(uint32_t (&r)[16], uint32_t (&q)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : TMEM_RW_REGS16(r), TMEM_RW_REGS16(q) :: "memory");
}
__device__ __forceinline__ void tmem_wait_ld1(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : TMEM_RW_REGS16(r) :: "memory");
}
__device__ __forceinline__ void tmem_wait_ld4(uint32_t (&r)[4]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]) :: "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ double u2d(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }


// S between tensor memory and a cache slot in global memory ([chunk][thread] double2: coalesced both ways).  Kept out of
// line: inlined into the solve loop they cost the hot path registers (config 2 ran 4 % slower).
template <int HALF, int NCT>
__device__ __noinline__ void tm_load_operator(uint32_t taddr, const double2 *src) {
  constexpr int SCHUNKS = HALF / 2;
#pragma unroll
  for (int g = 0; g < (2 * HALF) / 16; ++g) {
    uint32_t w[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double2 v = __ldcs(src + (size_t)(4 * g + j) * NCT);
      w[4 * j] = (uint32_t)__double2loint(v.x); w[4 * j + 1] = (uint32_t)__double2hiint(v.x);
      w[4 * j + 2] = (uint32_t)__double2loint(v.y); w[4 * j + 3] = (uint32_t)__double2hiint(v.y);
    }
    tmem_st16(taddr + 16 * g, w);
  }
#pragma unroll
  for (int c = 4 * ((2 * HALF) / 16); c < SCHUNKS; ++c) {
    const double2 v = __ldcs(src + (size_t)c * NCT);
    const uint32_t w4[4] = {(uint32_t)__double2loint(v.x), (uint32_t)__double2hiint(v.x), (uint32_t)__double2loint(v.y),
                            (uint32_t)__double2hiint(v.y)};
    tmem_st4(taddr + 4 * c, w4);
  }
  tmem_wait_st();
}
template <int HALF, int NCT>
__device__ __noinline__ void tm_store_operator(uint32_t taddr, double2 *dst) {
  constexpr int SCHUNKS = HALF / 2;
#pragma unroll
  for (int g = 0; g < (2 * HALF) / 16; ++g) {
    uint32_t w[16];
    tmem_ld16(taddr + 16 * g, w);
    tmem_wait_ld1(w);
#pragma unroll
    for (int j = 0; j < 4; ++j) __stcs(dst + (size_t)(4 * g + j) * NCT, make_double2(u2d(w[4 * j], w[4 * j + 1]), u2d(w[4 * j + 2], w[4 * j + 3])));
  }
#pragma unroll
  for (int c = 4 * ((2 * HALF) / 16); c < SCHUNKS; ++c) {
    uint32_t w4[4];
    tmem_ld4(taddr + 4 * c, w4);
    tmem_wait_ld4(w4);
    __stcs(dst + (size_t)c * NCT, make_double2(u2d(w4[0], w4[1]), u2d(w4[2], w4[3])));
  }
}

// Row re-typing (OSQP auxil.c update_rho_vec; reference src/trajectorySimulate.py:342 -> prob.update(l, u)): when the scaled
// bounds of a row come within RHO_TOL of each other OSQP re-classifies it as an equality, rho_vec[i] becomes 1e3 rho and the
// KKT matrix is refactored.  Only the velocity 1-norm rows of the LOS blocks (row nX + 5k + 3, k <= Nb: upper bound
// |p^ - r|_1) can do that, once a lane parks within centimetres of the target.  Such a row a has +-E_r D_c at the two
// velocity columns of stage k and (k < Nc) the soft-constraint entry at the stage's slack column, so M(rho) gains
// (1e3 - 1) rho a a' per re-typed row and the operator in tensor memory is corrected by one Sherman-Morrison step per
// row:  S <- S - (S a)(S a)' / (1 / ((1e3 - 1) rho) + a' S a).
// Every thread of the team calls this (it synchronises); wbuf is an N-vector scratch in shared memory (padding zero).
template <int HALF, int NCT>
__device__ __noinline__ void tm_retype_operator(uint32_t taddr, double *wbuf, const double *Ev, const double *Dv, int nX,
                                                unsigned mask, int variant, double rho, int tid, int col, bool has_col,
                                                const int *r3c, const double *r3v) {
  const int half = tid & 1;
  const bool col_warp = tid < NCT;
  const double delta_inv = 1.0 / ((MPCB_RHO_EQ - 1.0) * rho);
  for (int k = 0; (mask >> k) != 0u; ++k) {
    if (!((mask >> k) & 1u)) continue;
    const int r = nX + 5 * k + 3, cA = 4 * k + 2;
    const double al = ((variant & 1) ? -1.0 : 1.0) * Ev[r] * Dv[cA], be = ((variant & 2) ? -1.0 : 1.0) * Ev[r] * Dv[cA + 1];
    const int hA = cA / HALF, off = cA - hA * HALF;          // cA is even and HALF is even: both columns sit in one half, 4-column aligned
    const int cS = r3c[k];
    const double ga = cS >= 0 ? r3v[k] : 0.0;
    const int hS = cS >= 0 ? cS / HALF : 0, offS = cS >= 0 ? cS - hS * HALF : 0;
    double w = 0.0;
    if (col_warp) {                                            // w = S a: entries cA, cA + 1, cS of every row
      uint32_t c4[4];
      tmem_ld4(taddr + 2 * off, c4);
      tmem_wait_ld4(c4);
      if (half == hA) w = al * u2d(c4[0], c4[1]) + be * u2d(c4[2], c4[3]);
      tmem_ld4(taddr + 2 * (offS & ~1), c4);
      tmem_wait_ld4(c4);
      if (half == hS) w = fma(ga, (offS & 1) ? u2d(c4[2], c4[3]) : u2d(c4[0], c4[1]), w);
      w += __shfl_xor_sync(0xffffffffu, w, 1);
      if (has_col && half == 0) wbuf[col] = w;
    }
    __syncthreads();
    const double f = -w / (delta_inv + al * wbuf[cA] + be * wbuf[cA + 1] + (cS >= 0 ? ga * wbuf[cS] : 0.0));
    if (col_warp) {
      const double *wc = wbuf + half * HALF;
#pragma unroll 1
      for (int c = 0; c < HALF; c += 2) {
        uint32_t c4[4];
        tmem_ld4(taddr + 2 * c, c4);
        tmem_wait_ld4(c4);
        const double s0 = fma(f, wc[c], u2d(c4[0], c4[1])), s1 = fma(f, wc[c + 1], u2d(c4[2], c4[3]));
        const uint32_t w4[4] = {(uint32_t)__double2loint(s0), (uint32_t)__double2hiint(s0), (uint32_t)__double2loint(s1),
                                (uint32_t)__double2hiint(s1)};
        tmem_st4(taddr + 2 * c, w4);
      }
      tmem_wait_st();
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// N variables, M rows (M <= 256, 2N <= 256); WA (<= 8), WAT2 (<= 8, entry PAIRS), WP are the maximum
// ELL widths the instantiation supports.  Of the HALF entries of S a thread owns, the last SS live in
// shared memory ([SS][NCT], conflict-free) and the rest in registers.
//
// TM (tensor-memory mode): the whole of S lives in TMEM (2*HALF 32-bit columns per thread, 256 columns per CTA, two
// CTAs per SM = all 512), read back chunk by chunk with tcgen05.ld in the mat-vec.  The LSU data pipe -- not FP64
// -- is what two teams per SM saturate (tools/ubench_tmem.cu, ubench_lds.cu: every LDS.64 a warp issues costs
// the pipe a cycle, broadcast or not), so in this mode the registers S vacated hold the thread's rows of A / A'
// (values and packed column indices), which removes the table loads from every iteration, and SS must be 0.
// WATX: pair-slots beyond the WAT2 register-resident ones, read from shared memory by the few warps whose columns are
// that long (Nx = 20 / 30: the two disturbance columns have Nx+1 entries).  TT threads per team, CT teams per SM.
template <int N, int M, int WA, int WAT2, int WATX, int WP, int SS, bool TM, int TT, int CT>
__global__ void __launch_bounds__(TT, CT) team_kernel(const __grid_constant__ TeamArgs a) {
  constexpr int TEAM = TT;
  constexpr int HALF = ((N + 3) / 4) * 2;  // entries of a row of S per thread (even); 2*HALF >= N
  constexpr int NP2 = 2 * HALF;            // padded n-vector length in shared memory / Vk row length
  constexpr int MP = (M + 1) & ~1;
  constexpr int NW = TEAM / 32;
  constexpr int NCT = ((2 * N + 31) / 32) * 32;   // threads that run the pair phases: whole warps (shuffles stay convergent)
  constexpr int SR = HALF - SS;            // entries of S per thread kept in registers
  static_assert(SS % 2 == 0 && SR % 2 == 0 && SR > 0, "S split must be even");
  constexpr int TNEED = ((TEAM + 127) / 128) * 2 * HALF;     // TMEM columns: (warps per lane quarter) x (columns per thread)
  constexpr int TALLOC = TNEED <= 128 ? 128 : (TNEED <= 256 ? 256 : 512);
  static_assert(!TM || (SS == 0 && HALF >= 16 && HALF % 2 == 0 && TNEED <= 512 && CT * TALLOC <= 512), "tensor-memory mode: S does not fit");
  static_assert(2 * N <= TEAM && M <= TEAM && WA <= 8 && WAT2 <= 8 && WAT2 + WATX <= 16 && (TM || WATX == 0) && TEAM <= 512, "team too small");
  extern __shared__ __align__(128) unsigned char smem[];
  const TeamHdr &h = a.hdr;
  const int tid = threadIdx.x, warp = tid >> 5, lid = tid & 31;

  // ---- shared memory carve-up: [blob | vbuf MP | rbuf NP2 | xtbuf NP2 | dk NP2 | red 16*NW | lo hi rinv dy MP each | Ssm SS*NCT | ukf | ctx]
  double *vbuf = reinterpret_cast<double *>(smem + h.total);
  double *rbuf = vbuf + MP;
  double *xtbuf = rbuf + NP2;
  double *dk = xtbuf + NP2;
  double *red = dk + NP2;
  double *lobuf = red + 16 * NW;
  double *hibuf = lobuf + MP;
  double *rinvbuf = hibuf + MP;
  double *dybuf = rinvbuf + MP;
  double *Ssm = dybuf + MP;
  UkfScratch &ukf = *reinterpret_cast<UkfScratch *>(Ssm + SS * NCT);
  LaneCtx &L = *reinterpret_cast<LaneCtx *>(reinterpret_cast<unsigned char *>(&ukf) + ((sizeof(UkfScratch) + 15) & ~15));
  __shared__ int s_lane;
  __shared__ unsigned s_eqmask;

  for (int o = tid * 16; o < h.total; o += TEAM * 16)
    *reinterpret_cast<int4 *>(smem + o) = *reinterpret_cast<const int4 *>(a.blob + o);
  if (tid < NP2 - N) { rbuf[N + tid] = 0.0; xtbuf[N + tid] = 0.0; dk[N + tid] = 0.0; }
  if (tid < 36) { ukf.Ao[tid] = a.sc.Ao[tid]; ukf.Qw[tid] = a.sc.Qw[tid]; }
  if (tid < 12) ukf.Bou[tid] = a.sc.Bou[tid];
  if (tid == 0) ukf.dbg = a.tot;
  for (int o = tid; o < SS * NCT; o += TEAM) Ssm[o] = 0.0;
  __shared__ uint32_t s_tmem;
  if (TM) {
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tmem)), "n"(TALLOC));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
  }
  __syncthreads();
  if (TM) asm volatile("tcgen05.fence::after_thread_sync;");
  // this thread's 2*HALF columns: lane quarter warp%4, the upper four warps sit beside the lower four
  const uint32_t taddr = TM ? s_tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(2 * HALF * (warp >> 2)) : 0u;

  const double *qv = reinterpret_cast<const double *>(smem + h.off_q);
  const double *Dv = reinterpret_cast<const double *>(smem + h.off_D);
  const double *Dinv = reinterpret_cast<const double *>(smem + h.off_Dinv);
  const double *Ev = reinterpret_cast<const double *>(smem + h.off_E);
  const double *Einv = reinterpret_cast<const double *>(smem + h.off_Einv);
  const double *lt = reinterpret_cast<const double *>(smem + h.off_lt);
  const double *ut = reinterpret_cast<const double *>(smem + h.off_ut);
  double *Avals = reinterpret_cast<double *>(smem + h.off_Av);
  double *ATvals = reinterpret_cast<double *>(smem + h.off_ATv);
  const double *Pvals = reinterpret_cast<const double *>(smem + h.off_Pv);
  const uint4 *Acols = reinterpret_cast<const uint4 *>(smem + h.off_Ac);
  const uint4 *ATcols = reinterpret_cast<const uint4 *>(smem + h.off_ATc);
  const uint16_t *Pcols = reinterpret_cast<const uint16_t *>(smem + h.off_Pc);
  const uint8_t *flags = reinterpret_cast<const uint8_t *>(smem + h.off_flags);
  const int4 *patch = reinterpret_cast<const int4 *>(smem + h.off_patch);

  // roles: thread pair (2p, 2p+1) <-> variable col = colmap[p] / row col of S; thread t <-> row rowmap[t] of A
  const int pairi = tid >> 1, half = tid & 1;
  const bool has_col = pairi < N;        // pair member
  const bool col_warp = tid < NCT;       // warp-uniform: executes the pair phases (S = 0 for pairs beyond N)
  const bool has_row = tid < M;
  const int col = has_col ? reinterpret_cast<const uint16_t *>(smem + h.off_colmap)[pairi] : 0;
  const int row = has_row ? reinterpret_cast<const uint16_t *>(smem + h.off_rowmap)[tid] : 0;
  const int wA = h.wA[warp], wAT2 = h.wAT2[warp];          // warp-uniform ELL widths
  const double sigma = a.sigma, alpha = a.alpha, oma = 1.0 - a.alpha;

  // A warp-uniform switch picks a fully unrolled body, so every load of a phase is in flight at once.
  // tensor-memory mode: this thread's row of A and half-column of A' in registers (reloaded when the sign variant changes)
  double Ar[TM ? WA : 1], ATr[TM ? WAT2 : 1];
  uint4 Acr = make_uint4(0, 0, 0, 0), ATcr = make_uint4(0, 0, 0, 0);
  auto load_tables = [&]() {
    if constexpr (TM) {
#pragma unroll
      for (int e = 0; e < WA; ++e) Ar[e] = has_row ? Avals[e * MP + tid] : 0.0;
#pragma unroll
      for (int e = 0; e < WAT2; ++e) ATr[e] = col_warp ? ATvals[e * NCT + tid] : 0.0;
      if (has_row) Acr = Acols[tid];
      if (col_warp) ATcr = ATcols[tid];
    }
  };
  load_tables();
  auto applyA = [&](const double *vec) -> double {          // row `row` of A times vec (thread tid < M)
    if constexpr (TM) {
      // values and indices are registers: padding a row to its class width costs one broadcast load and one FMA per
      // pad, a jump table (LDC + BRX, on the critical path of every iteration) costs more
      if (wA > 5) return ell_dot_reg<WA>(Ar, Acr, vec);
      if (wA > 2) return ell_dot_reg<(WA >= 5 ? 5 : WA)>(Ar, Acr, vec);
      return ell_dot_reg<(WA >= 2 ? 2 : WA)>(Ar, Acr, vec);
    }
    const uint4 c = Acols[tid];
    const double *v = Avals + tid;
    switch (wA) {
      case 1: return ell_dot<1>(v, MP, c, vec);
      case 2: return ell_dot<2>(v, MP, c, vec);
      case 3: return ell_dot<3>(v, MP, c, vec);
      case 4: return ell_dot<4>(v, MP, c, vec);
      case 5: return ell_dot<5>(v, MP, c, vec);
      case 6: return ell_dot<6>(v, MP, c, vec);
      case 7: return ell_dot<7>(v, MP, c, vec);
      case 8: return ell_dot<8>(v, MP, c, vec);
      default: return 0.0;
    }
  };
  auto applyAT = [&](const double *vec) -> double {         // (A' vec)[col], pair-summed (threads < NCT)
    const uint4 c = ATcols[tid];
    const double *v = ATvals + tid;
    double acc;
    if constexpr (TM) {
      if (wAT2 > 4) acc = ell_dot_reg<WAT2>(ATr, ATcr, vec);
      else if (wAT2 > 2) acc = ell_dot_reg<(WAT2 >= 4 ? 4 : WAT2)>(ATr, ATcr, vec);
      else acc = ell_dot_reg<(WAT2 >= 2 ? 2 : WAT2)>(ATr, ATcr, vec);
      if constexpr (WATX > 0) {
        if (wAT2 > WAT2) {               // warp-uniform: the warp that owns the long columns
          const uint4 c2 = reinterpret_cast<const uint4 *>(smem + h.off_ATc2)[tid];
          const unsigned cw[8] = {ATcr.x, ATcr.y, ATcr.z, ATcr.w, c2.x, c2.y, c2.z, c2.w};
          double acc1 = 0.0;
#pragma unroll
          for (int e = WAT2; e < WAT2 + WATX; ++e)
            if (e < wAT2) {
              const unsigned idx = (e & 1) ? (cw[e >> 1] >> 16) : (cw[e >> 1] & 0xffffu);
              if (e & 1) acc1 = fma(ATvals[e * NCT + tid], vec[idx], acc1);
              else acc = fma(ATvals[e * NCT + tid], vec[idx], acc);
            }
          acc += acc1;
        }
      }
      return acc + __shfl_xor_sync(0xffffffffu, acc, 1);
    }
    switch (wAT2) {
      case 1: acc = ell_dot<1>(v, NCT, c, vec); break;
      case 2: acc = ell_dot<2>(v, NCT, c, vec); break;
      case 3: acc = ell_dot<3>(v, NCT, c, vec); break;
      case 4: acc = ell_dot<4>(v, NCT, c, vec); break;
      case 5: acc = ell_dot<5>(v, NCT, c, vec); break;
      case 6: acc = ell_dot<6>(v, NCT, c, vec); break;
      case 7: acc = ell_dot<7>(v, NCT, c, vec); break;
      case 8: acc = ell_dot<8>(v, NCT, c, vec); break;
      default: acc = 0.0; break;
    }
    return acc + __shfl_xor_sync(0xffffffffu, acc, 1);
  };
  auto applyP = [&](const double *vec) -> double {          // row col of P times vec (both pair members)
    double acc = 0.0;
#pragma unroll
    for (int e = 0; e < WP; ++e) acc = fma(Pvals[e * (NCT / 2) + pairi], vec[Pcols[e * (NCT / 2) + pairi]], acc);
    return acc;
  };
  auto team_max = [&](auto &vals) {                  // vals: double[CNT], CNT <= 12, stays in registers
    constexpr int CNT = sizeof(vals) / sizeof(double);
#pragma unroll
    for (int q = 0; q < CNT; ++q) vals[q] = warp_max_nonneg(vals[q]);    // every reduced quantity is an absolute value
    __syncthreads();
    if (lid == 0) {
#pragma unroll
      for (int q = 0; q < CNT; ++q) red[warp * 16 + q] = vals[q];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < CNT; ++q) {
      double v = red[q];
#pragma unroll
      for (int w = 1; w < NW; ++w) v = fmax(v, red[w * 16 + q]);
      vals[q] = v;
    }
  };

  double S[SR];                          // S[col][half*HALF .. half*HALF+SR); the next SS entries are in Ssm[.][tid]
  int op_variant = 0;                    // sign variant baked into Avals / ATvals (blob = variant 0)
  double op_rho = -1.0;                  // (rho, variant) the operator S currently held was built for: S depends on nothing
  int s_variant = -1;                    // else, so it survives from one lane to the next when both match
  unsigned op_mask = 0u;                 // re-typed rows folded into S (tm_retype_operator); the operator caches hold mask-0 operators only
  constexpr int SCHUNKS = HALF / 2;      // double2 chunks of S per thread (operator cache layout [chunk][thread])
  auto s_load = [&](const double2 *src) { if constexpr (TM) tm_load_operator<HALF, NCT>(taddr, src); };     // col_warp threads,
  auto s_store = [&](double2 *dst) { if constexpr (TM) tm_store_operator<HALF, NCT>(taddr, dst); };         // pointer offset by tid
  // Compiled in for the larger families only: at n = 81 a rebuild costs about one solve, flips at unchanged rho are 5 % of
  // config 2's rebuilds, and the extra code cost the hot loop 2 %.
  constexpr bool OC = TM && N > 100;
  __shared__ double s_oc_rho[TEAM_OPC];      // tags of the team's operator cache (rho < 0: empty), written by thread 0
  __shared__ int s_oc_var[TEAM_OPC], s_oc_next;
  if (tid < TEAM_OPC) { s_oc_rho[tid] = -1.0; s_oc_var[tid] = -1; }
  if (tid == 0) s_oc_next = 0;
  unsigned long long my_iters = 0, my_rebuilds = 0;
  TP_DECL

  while (true) {
    __syncthreads();
    if (tid == 0) {
      int q = atomicAdd(a.queue, 1);
      if (a.list) {                      // q-th entry of the concatenated per-variant lists
        int v = 0;
        for (; v < 4; ++v) {
          const int c = a.cnt[v];
          if (q < c) break;
          q -= c;
        }
        if (v < 4) {
          q = a.list[(size_t)v * a.B + q];
          L.variant = v;
        } else {
          q = a.B;
        }
      }
      s_lane = q;
    }
    __syncthreads();
    const int ln = s_lane;
    if (ln >= a.B) break;
    if (tid == 0) lane_init(a, L, ln);
    bool first_solve = true;             // list / resume modes: the lane's first solve here may already be under way
    // solver iterates: x[col] (both pair members), z[row] / y[row]
    double x = 0.0, z = 0.0, y = 0.0;
    if (a.warm) {
      if (has_col) x = a.xs[(size_t)ln * N + col];
      if (has_row) {
        z = a.zs[(size_t)ln * M + row];
        y = a.ys[(size_t)ln * M + row];
      }
    }
    __syncthreads();
    if (TM && a.scache && (L.rho != op_rho || L.variant != s_variant) && L.c_rho == L.rho && L.c_var == L.variant) {
      // this lane's operator is in its cache slot: HBM -> registers -> tensor memory
      if (col_warp) s_load(reinterpret_cast<const double2 *>(a.scache) + (size_t)ln * (SCHUNKS * NCT) + tid);
      op_rho = L.rho;
      s_variant = L.variant;
      op_mask = 0u;
    }
    TP_MARK(4)

    // =============================== control steps ===============================
    while (!L.fin) {
      const int variant = L.variant;
      double rho = L.rho;
      int r3k = -1;                      // >= 0: this thread's row is the velocity-bound row of stage r3k
      if (has_row) {
        double lo = lt[row], hi = ut[row];
        if (row < 4) lo = hi = -L.par[row] * Ev[row];
        else if (row >= M - 2) lo = hi = L.par[5 + (row - (M - 2))] * Ev[row];
        else if (row >= a.nX && row < a.nX + 5 * (a.Nb + 1) && (row - a.nX) % 5 == 3) {
          hi = L.par[4] * Ev[row];
          r3k = (row - a.nX) / 5;
        }
        lobuf[tid] = lo;                               // own slot only: no barrier needed before own reads
        hibuf[tid] = hi;
      }
      if (tid == 0) {                                  // rows OSQP treats as equalities at this step (tm_retype_operator)
        unsigned mk = 0u;
        for (int k = 0; k <= a.Nb && k < 32; ++k) {
          const int r = a.nX + 5 * k + 3;
          if (L.par[4] * Ev[r] - lt[r] < MPCB_RHO_TOL) mk |= 1u << k;
        }
        if (mk) L.flip = 1;
        s_eqmask = mk;
      }
      if (variant != op_variant) {       // re-sign the (Nx+1) velocity rows of A and A'
        for (int p = tid; p < h.n_patch; p += TEAM) {
          const int4 e = patch[p];
          const bool neg_new = (e.z == 1) ? (variant & 1) : ((variant & 2) != 0);
          const bool neg_old = (e.z == 1) ? (op_variant & 1) : ((op_variant & 2) != 0);
          if (neg_new != neg_old) {
            Avals[e.x] = -Avals[e.x];
            ATvals[e.y] = -ATvals[e.y];
          }
        }
      }
      int iter = (a.list && first_solve) ? L.iter : 0, st = -10;        // list mode: a solve may span several visits
      first_solve = false;
      int visit_left = a.visit_iters;
      const bool resigned = variant != op_variant;
      op_variant = variant;
      __syncthreads();
      const unsigned eqmask = (TM && h.r3ok) ? s_eqmask : 0u;      // register-resident operator (MPCB_TEAM=regs): re-typing is only counted
      bool need_op = (variant != s_variant) || (rho != op_rho) || (eqmask != op_mask);
      uint8_t fl = has_row ? flags[row] : (uint8_t)8;
      if (r3k >= 0 && ((eqmask >> r3k) & 1u)) fl |= 4;  // re-typed row: rho_vec = 1e3 rho
      if (TM && resigned) load_tables();

      // =============================== one solve ===============================
      while (st == -10) {
        int oc_hit = -1;
        if (OC && need_op && a.ocache) {
          // A sign-variant flip at unchanged rho tends to flip back (95 % of config 4's rebuilds): park the OUTGOING
          // operator before it is overwritten, unless the cache has it already.  Operators replaced because rho adapted are
          // not parked (they rarely return, and config 2 would pay a 64 KB store per rebuild for a 5 % hit rate).
          if (op_rho == rho && s_variant != variant && s_variant >= 0 && op_mask == 0u) {
            bool have = false;
#pragma unroll
            for (int q = 0; q < TEAM_OPC; ++q) have = have || (s_oc_rho[q] == op_rho && s_oc_var[q] == s_variant);
            const int victim = s_oc_next;
            __syncthreads();
            if (!have) {
              if (col_warp)
                s_store(reinterpret_cast<double2 *>(a.ocache) + ((size_t)blockIdx.x * TEAM_OPC + victim) * (SCHUNKS * NCT) + tid);
              if (tid == 0) {
                s_oc_rho[victim] = op_rho;
                s_oc_var[victim] = s_variant;
                s_oc_next = (victim + 1) % TEAM_OPC;
              }
            }
            __syncthreads();
          }
#pragma unroll
          for (int q = 0; q < TEAM_OPC; ++q)
            if (s_oc_rho[q] == rho && s_oc_var[q] == variant) oc_hit = q;
        }
        if (need_op && oc_hit >= 0) {    // one of the team's last operators: reload it (bit-identical to a rebuild)
          if (col_warp)
            s_load(reinterpret_cast<const double2 *>(a.ocache) + ((size_t)blockIdx.x * TEAM_OPC + oc_hit) * (SCHUNKS * NCT) + tid);
          op_rho = rho;
          s_variant = variant;
          op_mask = 0u;
          need_op = false;
        }
        if (need_op) {                   // S = V diag(1/(1+rho*lam)) V'
          if (tid < N) dk[tid] = 1.0 / (1.0 + rho * a.lam[variant * N + tid]);
          __syncthreads();
          if constexpr (TM) {
            // passes over V, 16 entries of this thread's half row at a time, keep the accumulators small; the summation
            // order over k is the register version's, so S is bit-identical to it
            if (col_warp) {
              const double *Vk = a.Vk[variant];
              // RB entries of the half row per pass: 16 at two teams per SM (128 registers), 32 with a whole SM's registers --
              // V (211 KB per variant at n = 161) comes from L2, so fewer, fatter passes mean more loads in flight
              constexpr int RB = (CT == 1) ? 32 : 16;
              constexpr int NPASS = (HALF + RB - 1) / RB;
#pragma unroll 1
              for (int pass = 0; pass < NPASS; ++pass) {
                double T[RB];
#pragma unroll
                for (int j = 0; j < RB; ++j) T[j] = 0.0;
                const int cnt2 = (pass < NPASS - 1) ? RB / 2 : (HALF - RB * (NPASS - 1)) / 2;
#pragma unroll 3
                for (int k = 0; k < N; ++k) {
                  const double *vrow = Vk + (size_t)k * NP2;
                  const double tk = has_col ? __ldg(vrow + col) * dk[k] : 0.0;
                  const double2 *r2 = reinterpret_cast<const double2 *>(vrow + half * HALF + RB * pass);
#pragma unroll
                  for (int j = 0; j < RB / 2; ++j)
                    if (j < cnt2) {
                      const double2 vv = __ldg(r2 + j);
                      T[2 * j] = fma(tk, vv.x, T[2 * j]);
                      T[2 * j + 1] = fma(tk, vv.y, T[2 * j + 1]);
                    }
                }
#pragma unroll
                for (int hh = 0; hh < RB / 8; ++hh) {
                  uint32_t w[16];
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    w[2 * j] = (uint32_t)__double2loint(T[8 * hh + j]);
                    w[2 * j + 1] = (uint32_t)__double2hiint(T[8 * hh + j]);
                  }
                  const int c0 = 2 * RB * pass + 16 * hh;            // first column of these 8 doubles
                  if (c0 + 16 <= 2 * HALF) tmem_st16(taddr + c0, w);
                  else if (c0 < 2 * HALF) {                          // tail: 2*HALF - c0 columns, in 4-column pieces
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                      if (c0 + 4 * q < 2 * HALF) {
                        const uint32_t w4[4] = {w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]};
                        tmem_st4(taddr + c0 + 4 * q, w4);
                      }
                  }
                }
              }
              tmem_wait_st();
            }
          } else {
#pragma unroll
          for (int j = 0; j < SR; ++j) S[j] = 0.0;
          if (has_col) {
            const double *Vk = a.Vk[variant];
            double T[SS > 0 ? SS : 1];
#pragma unroll
            for (int j = 0; j < SS; ++j) T[j] = 0.0;
            for (int k = 0; k < N; ++k) {
              const double *vrow = Vk + (size_t)k * NP2;
              const double tk = __ldg(vrow + col) * dk[k];
              const double2 *r2 = reinterpret_cast<const double2 *>(vrow + half * HALF);
#pragma unroll
              for (int j = 0; j < SR / 2; ++j) {
                const double2 vv = __ldg(r2 + j);
                S[2 * j] = fma(tk, vv.x, S[2 * j]);
                S[2 * j + 1] = fma(tk, vv.y, S[2 * j + 1]);
              }
#pragma unroll
              for (int j = 0; j < SS / 2; ++j) {
                const double2 vv = __ldg(r2 + SR / 2 + j);
                T[2 * j] = fma(tk, vv.x, T[2 * j]);
                T[2 * j + 1] = fma(tk, vv.y, T[2 * j + 1]);
              }
            }
#pragma unroll
            for (int j = 0; j < SS; ++j) Ssm[j * NCT + tid] = T[j];
          }
          }
          op_rho = rho;
          s_variant = variant;
          op_mask = 0u;
          need_op = false;
          ++my_rebuilds;
          TP_MARK(2)
        }
        if constexpr (TM) {
          if (op_mask != eqmask) {       // (team-uniform) fold the re-typed rows into the fresh S(rho)
            __syncthreads();
            tm_retype_operator<HALF, NCT>(taddr, xtbuf, Ev, Dv, a.nX, eqmask, variant, rho, tid, col, has_col, h.r3c, h.r3v);
            op_mask = eqmask;
          }
        }
        const double rv = (fl & 8) ? MPCB_RHO_MIN : ((fl & 4) ? MPCB_RHO_EQ * rho : rho);
        if (has_row) {
          rinvbuf[tid] = 1.0 / rv;
          vbuf[row] = rv * z - y;
        }
        __syncthreads();
        // ---- check_every ADMM iterations
        const int last_it = a.check_every - 1;
        for (int it = 0; it <= last_it; ++it) {
          if (col_warp) {
            const double s = applyAT(vbuf);
            if (has_col && half == 0) rbuf[col] = sigma * x - qv[col] + s;
          }
          __syncthreads();
          if (TM && col_warp) {
            // S streams out of tensor memory in 16-column chunks (8 doubles), two chunks in flight
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            const double2 *r2 = reinterpret_cast<const double2 *>(rbuf + half * HALF);
#if TEAM_TM_CHUNK == 8
            uint32_t ca[8], cb[8];                       // 8-column chunks (4 doubles): half the staging registers
            auto use = [&](const uint32_t (&c)[8], int q) {
              const double2 r0 = r2[2 * q], r1 = r2[2 * q + 1];
              a0 = fma(u2d(c[0], c[1]), r0.x, a0);
              a1 = fma(u2d(c[2], c[3]), r0.y, a1);
              a2 = fma(u2d(c[4], c[5]), r1.x, a2);
              a3 = fma(u2d(c[6], c[7]), r1.y, a3);
            };
            constexpr int NCH = (2 * HALF) / 8, CW = 8;
            tmem_ld8(taddr, ca);
            tmem_ld8(taddr + 8, cb);
            tmem_wait_ld8x2(ca, cb);
#pragma unroll
            for (int q = 0; q < NCH; ++q) {
              if (q & 1) {
                use(cb, q);
                if (q + 2 < NCH) tmem_ld8(taddr + 8 * (q + 2), cb);
              } else {
                use(ca, q);
                if (q + 2 < NCH) tmem_ld8(taddr + 8 * (q + 2), ca);
              }
              if ((q & 1) && q + 1 < NCH) tmem_wait_ld8x2(ca, cb);
              if (!(q & 1) && q + 1 < NCH && q + 2 >= NCH) tmem_wait_ld8x2(ca, cb);
            }
#else
            uint32_t ca[16], cb[16];
            auto use = [&](const uint32_t (&c)[16], int q) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const double2 rr = r2[4 * q + j];
                const double s0 = u2d(c[4 * j], c[4 * j + 1]), s1 = u2d(c[4 * j + 2], c[4 * j + 3]);
                if (j & 1) {
                  a2 = fma(s0, rr.x, a2);
                  a3 = fma(s1, rr.y, a3);
                } else {
                  a0 = fma(s0, rr.x, a0);
                  a1 = fma(s1, rr.y, a1);
                }
              }
            };
            constexpr int NCH = (2 * HALF) / 16, CW = 16;       // full chunks; the remaining (2*HALF - 16*NCH) / 4 doubles pairs come as x4 loads
            tmem_ld16(taddr, ca);
            tmem_ld16(taddr + 16, cb);
            tmem_wait_ld2(ca, cb);
#pragma unroll
            for (int q = 0; q < NCH; ++q) {
              if (q & 1) {
                use(cb, q);
                if (q + 2 < NCH) tmem_ld16(taddr + 16 * (q + 2), cb);
              } else {
                use(ca, q);
                if (q + 2 < NCH) tmem_ld16(taddr + 16 * (q + 2), ca);
              }
              if ((q & 1) && q + 1 < NCH) tmem_wait_ld2(ca, cb);
              if (!(q & 1) && q + 1 < NCH && q + 2 >= NCH) tmem_wait_ld2(ca, cb);
            }
#endif
#pragma unroll
            for (int c0 = CW * NCH; c0 < 2 * HALF; c0 += 4) {
              uint32_t c4[4];
              tmem_ld4(taddr + c0, c4);
              tmem_wait_ld4(c4);
              const double2 rr = r2[c0 / 4];
              a0 = fma(u2d(c4[0], c4[1]), rr.x, a0);
              a1 = fma(u2d(c4[2], c4[3]), rr.y, a1);
            }
            double xt = (a0 + a1) + (a2 + a3);
            xt += __shfl_xor_sync(0xffffffffu, xt, 1);
            if (has_col && half == 0) xtbuf[col] = xt;
            x = alpha * xt + oma * x;
          }
          if (!TM && col_warp) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            const double2 *r2 = reinterpret_cast<const double2 *>(rbuf + half * HALF);
#pragma unroll
            for (int j = 0; j < SS / 2; ++j) {          // shared-memory part first: its loads have the longest chain
              const double2 rr = r2[SR / 2 + j];
              const double s0 = Ssm[(2 * j) * NCT + tid], s1 = Ssm[(2 * j + 1) * NCT + tid];
              if (j & 1) {
                a2 = fma(s0, rr.x, a2);
                a3 = fma(s1, rr.y, a3);
              } else {
                a0 = fma(s0, rr.x, a0);
                a1 = fma(s1, rr.y, a1);
              }
            }
#pragma unroll
            for (int j = 0; j < SR / 2; ++j) {
              const double2 rr = r2[j];
              if (j & 1) {
                a2 = fma(S[2 * j], rr.x, a2);
                a3 = fma(S[2 * j + 1], rr.y, a3);
              } else {
                a0 = fma(S[2 * j], rr.x, a0);
                a1 = fma(S[2 * j + 1], rr.y, a1);
              }
            }
            double xt = (a0 + a1) + (a2 + a3);
            xt += __shfl_xor_sync(0xffffffffu, xt, 1);
            if (has_col && half == 0) xtbuf[col] = xt;
            x = alpha * xt + oma * x;
          }
          __syncthreads();
          if (has_row) {
            const double zt = applyA(xtbuf);
            const double zr = alpha * zt + oma * z;
            const double zn = fmin(fmax(zr + rinvbuf[tid] * y, lobuf[tid]), hibuf[tid]);
            const double dy = rv * (zr - zn);
            y += dy;
            z = zn;
            vbuf[row] = rv * zn - y;
            if (it == last_it) {                       // only the block's last delta_y feeds the infeasibility test:
              double d = dy;                           // store it already projected (is_primal_infeasible, OSQP auxil.c)
              if ((fl & 3) == 3) d = 0.0;
              else if (fl & 2) d = fmin(d, 0.0);
              else if (fl & 1) d = fmax(d, 0.0);
              dybuf[row] = d;
            }
          }
          __syncthreads();
        }
        iter += a.check_every;
        my_iters += (unsigned long long)a.check_every;
        TP_MARK(0)

        // ---- update_info (OSQP auxil.c): unscaled residuals decide termination; the scaled ones are
        //      only needed when rho may adapt, so they are reduced lazily
        if (has_col && half == 0) xtbuf[col] = x;
        if (has_row) vbuf[row] = y;
        __syncthreads();
        // One reduction round serves both the termination test and the primal-infeasibility certificate
        // (||E dy||, ||D^-1 A'dy||, u'dy+ + l'dy-): the certificate's extra A' product is cheaper than the
        // two barriers a lazy evaluation costs, and most config-2 solves need it.
        double Ax = 0.0, Px = 0.0, dv = 0.0, l1 = 0.0;
        double mu[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) mu[q] = 0.0;
        if (has_row) {
          Ax = applyA(xtbuf);
          const double ei = Einv[row], d = dybuf[row];
          mu[0] = fabs(ei * (Ax - z));
          mu[1] = fabs(ei * z);
          mu[2] = fabs(ei * Ax);
          mu[6] = fabs(Ev[row] * d);
          l1 = hibuf[tid] * fmax(d, 0.0) + lobuf[tid] * fmin(d, 0.0);
        }
        const double Aty = col_warp ? applyAT(vbuf) : 0.0;
        const double Atd = col_warp ? applyAT(dybuf) : 0.0;
        if (has_col) {
          Px = applyP(xtbuf);
          dv = qv[col] + Px + Aty;
          const double di = Dinv[col];
          mu[3] = fabs(di * dv);
          mu[4] = fabs(di * Px);
          mu[5] = fabs(di * Aty);
          mu[7] = fabs(di * Atd);
        }
        l1 = warp_sum(l1);
        if (lid == 0) red[warp * 16 + 15] = l1;      // slot 15 is not touched by team_max (<= 12 values); its barriers publish it
        team_max(mu);
        const double pri_u = mu[0], nz_u = mu[1], nax_u = mu[2], dua_u = mu[3] * a.cinv, npx_u = mu[4], naty_u = mu[5];
        const double ndy = mu[6], natdy = mu[7];
        double lhs = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) lhs += red[w * 16 + 15];

        auto check = [&](double k) -> int {
          const double eps_p = k * a.eps_abs + k * a.eps_rel * fmax(nz_u, nax_u);
          const double eps_d = k * a.eps_abs + k * a.eps_rel * a.cinv * fmax(a.qn_unscaled, fmax(naty_u, npx_u));
          const bool prim_ok = pri_u < eps_p, dual_ok = dua_u < eps_d;
          if (prim_ok && dual_ok) return (k > 1.0) ? 2 : 1;
          if (!prim_ok) {
            const double eps_i = k * a.eps_pinf;
            if (ndy > MPCB_DIV_TOL && lhs < -eps_i * ndy && natdy < eps_i * ndy) return (k > 1.0) ? 3 : -3;
          }
          return -10;
        };
        st = check(1.0);
        if (st == -10) {
          if (a.adaptive && (iter % a.adapt_interval == 0)) {   // compute_rho_estimate on the SCALED residuals (OSQP 0.6.x)
            double ms[6];
            ms[0] = has_row ? fabs(Ax - z) : 0.0;
            ms[1] = has_row ? fabs(z) : 0.0;
            ms[2] = has_row ? fabs(Ax) : 0.0;
            ms[3] = has_col ? fabs(dv) : 0.0;
            ms[4] = has_col ? fabs(Px) : 0.0;
            ms[5] = has_col ? fabs(Aty) : 0.0;
            team_max(ms);
            const double pr = ms[0] / (fmax(ms[1], ms[2]) + 1e-10);
            const double du = ms[3] / (fmax(a.qn_scaled, fmax(ms[5], ms[4])) + 1e-10);
            double est = rho * sqrt(pr / (du + 1e-10));
            est = fmin(fmax(est, MPCB_RHO_MIN), MPCB_RHO_MAX);
            if (est > rho * a.adapt_tol || est < rho / a.adapt_tol) {
              rho = est;
              need_op = true;
            }
          }
          if (iter >= a.max_iter) {
            st = check(10.0);
            if (st == -10) st = -2;
          }
        }
        __syncthreads();                 // red / vbuf reads of this check are done before the next block writes
        TP_MARK(1)
        if (a.visit_iters > 0 && (visit_left -= a.check_every) <= 0) break;     // visit budget spent: the lane is re-listed
      }  // solve

      // ---- hand the result to the lane context, run the rest of the control step on warp 0
      if (tid == 0) {
        L.rho = rho;
        L.status = st;
        L.iter = iter;
        L.u0[0] = Dv[a.uoff] * xtbuf[a.uoff];
        L.u0[1] = Dv[a.uoff + 1] * xtbuf[a.uoff + 1];
        if (a.mode == MODE_QP_ONLY) L.fin = 1;
      }
      if (warp == 0 && a.mode != MODE_QP_ONLY) {
        __syncwarp();
        lane_post_step(a, L, ukf, lid);
      }
      __syncthreads();
      TP_MARK(3)
    }  // control steps

    // ---- list mode: leave the operator in the lane's cache slot unless the slot already holds it
    if constexpr (TM) {
      const bool put = a.list && a.scache && op_rho >= 0.0 && op_mask == 0u && (L.c_rho != op_rho || L.c_var != s_variant);
      __syncthreads();
      if (put) {
        if (col_warp) s_store(reinterpret_cast<double2 *>(a.scache) + (size_t)ln * (SCHUNKS * NCT) + tid);
        if (tid == 0) { L.c_rho = op_rho; L.c_var = s_variant; }
      }
    }
    // ---- lane done: persist solver iterates (QP seam warm start / mpcb_qp_get_state), final results
    if (has_col && half == 0) a.xs[(size_t)ln * N + col] = x;
    if (has_row) {
      a.zs[(size_t)ln * M + row] = z;
      a.ys[(size_t)ln * M + row] = y;
    }
    if (tid == 0) lane_finalize(a, L);
  }
  TP_FLUSH
  if (TM) {
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(TALLOC));
  }
  if (tid == 0) {
    if (my_iters) atomicAdd(&a.tot[0], my_iters);
    if (my_rebuilds) atomicAdd(&a.tot[2], my_rebuilds);
  }
}
