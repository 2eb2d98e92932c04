// ADMM tile kernel: the throughput-oriented solver block for LARGE batches.  Same contract as
// admm_block_kernel (admm.cuh): one check_termination period (25 iterations) of OSQP's ADMM for every
// live lane of this round, then termination / infeasibility / adaptive-rho logic -- but the lanes are
// tiled 8 per warp and the two dense mat-vecs of the spectral form
//     w  = diag(1/(1+rho*lam)) V' r        x~ = V w
// are batched over the tile on the FP64 tensor cores: mma.sync.aligned.m8n8k4 (DMMA), A operand = the 8
// lanes' vectors (8 x 4 slice), B operand = a 4 x 8 slice of V from shared memory, accumulators = an
// 8-lane x 8-column tile.  Every element of V loaded from shared memory feeds 8 lanes, which is what the
// warp-per-lane kernel (1 134 shared wavefronts per lane-iteration) and the register-resident team kernel
// (two lanes per SM) cannot offer; rho stays per lane (it only enters through the diagonal scaling).
// Everything is warp-local: a tile never waits on another warp (only __syncwarp between phases).
//
// Thread t of a warp: g = t >> 2 = lane-in-tile ("trajectory"), c = t & 3.
//   n-vectors in C-fragment layout: thread (g, c) owns entries 8*jt + 2c + {0,1} of trajectory g (x lives there);
//   sparse / elementwise work: thread (g, c) owns columns j = c (mod 4) and rows i = c (mod 4) of trajectory g.
#pragma once
#include "common.cuh"

struct TileHdr {
  int off_V;                                   // [NT8*8 rows][LDV] doubles, zero padded
  int off_lam, off_q, off_D, off_Dinv;         // [NPAD]
  int off_E, off_Einv, off_lt, off_ut;         // [MPAD]
  int off_Av, off_ATv, off_Pv;                 // ELL values [e][MPAD] / [e][NPAD] / [e][NPAD]
  int off_Ac, off_ATc, off_Pc;                 // uint8 column / row indices, same shapes
  int off_An, off_ATn, off_Pn;                 // uint8 entries per row / column
  int off_flags, off_pcode;                    // uint8 per row: rho class bits; bound patch code
  int total;
};

struct TileArgs {
  TileHdr hdr;
  const unsigned char *blob[4];                // one blob per velocity-sign variant
  int n, m, uoff, B;
  double sigma, alpha, eps_abs, eps_rel, eps_pinf, adapt_tol, cinv, qn_unscaled, qn_scaled;
  int check_every, adaptive, adapt_interval, max_iter;
  const int *cnt;        // [4] lanes per variant this round
  const int *list;       // [4][B]
  double *xs, *zs, *ys;  // [B][n], [B][m], [B][m]
  double *rho;
  int *iter, *status;
  const double *par;     // [7][B]
  double *u0;            // [2][B]
  uint8_t *lane_state;
  int *flip;
  unsigned long long *iter_total;
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double quad_max(double v) {
  v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmax(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ double quad_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// N variables, M rows, WARPS tiles per CTA.  WAM / WATM / WPM: maximum ELL widths (rows of A, columns of A, rows of P).
template <int N, int M, int WARPS, int WAM, int WATM, int WPM>
__global__ void __launch_bounds__(32 * WARPS, 1) admm_tile_kernel(const __grid_constant__ TileArgs a) {
  constexpr int KS = (N + 3) / 4;              // k-steps of 4
  constexpr int NT8 = (N + 7) / 8;             // 8-column output tiles
  constexpr int LDV = ((4 * KS - 4 + 15) / 16) * 16 + 4;   // smallest stride >= 4*KS that is 4 (mod 16): the "8 rows x 4 consecutive" fragment pattern is conflict free
  constexpr int LDN = LDV;                     // per-trajectory n-vector stride (same pattern)
  constexpr int LDM = ((M - 4 + 15) / 16) * 16 + 4;
  constexpr int NPAD = LDV, MPAD = LDM;
  constexpr int RN = (N + 3) / 4;              // columns per thread in the sparse phases
  constexpr int RM = (M + 3) / 4;              // rows per thread
  static_assert(4 * KS <= LDV && 8 * NT8 <= LDV + 4, "padding");
  extern __shared__ __align__(128) unsigned char smem[];
  const TileHdr &h = a.hdr;
  const int warp = threadIdx.x >> 5, lid = threadIdx.x & 31, g = lid >> 2, c = lid & 3;

  // ---- which (variant, group of WARPS tiles) is this CTA
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
  __syncthreads();

  const double *V = reinterpret_cast<const double *>(smem + h.off_V);
  const double *lam = reinterpret_cast<const double *>(smem + h.off_lam);
  const double *qv = reinterpret_cast<const double *>(smem + h.off_q);
  const double *Dv = reinterpret_cast<const double *>(smem + h.off_D);
  const double *Dinv = reinterpret_cast<const double *>(smem + h.off_Dinv);
  const double *Ev = reinterpret_cast<const double *>(smem + h.off_E);
  const double *Einv = reinterpret_cast<const double *>(smem + h.off_Einv);
  const double *lt = reinterpret_cast<const double *>(smem + h.off_lt);
  const double *ut = reinterpret_cast<const double *>(smem + h.off_ut);
  const double *Av = reinterpret_cast<const double *>(smem + h.off_Av);
  const double *ATv = reinterpret_cast<const double *>(smem + h.off_ATv);
  const double *Pv = reinterpret_cast<const double *>(smem + h.off_Pv);
  const uint8_t *Ac = smem + h.off_Ac, *ATc = smem + h.off_ATc, *Pc = smem + h.off_Pc;
  const uint8_t *An = smem + h.off_An, *ATn = smem + h.off_ATn, *Pn = smem + h.off_Pn;
  const uint8_t *flags = smem + h.off_flags, *pcode = smem + h.off_pcode;

  // per-warp buffers: nb (n-vector: r, then w, then x~), zb, yb, db (m-vectors), one row per trajectory
  double *wbase = reinterpret_cast<double *>(smem + h.total) + (size_t)warp * 8 * (LDN + 3 * LDM);
  double *nb = wbase + g * LDN;
  double *zb = wbase + 8 * LDN + g * LDM;
  double *yb = wbase + 8 * LDN + 8 * LDM + g * LDM;
  double *db = wbase + 8 * LDN + 16 * LDM + g * LDM;

  const int pos = (b * WARPS + warp) * 8 + g;
  const bool valid = pos < cnt_v;
  if (__ballot_sync(0xffffffffu, valid) == 0) return;           // whole tile empty
  const int ln = valid ? a.list[(size_t)v * a.B + pos] : 0;
  const size_t Bz = a.B;
  double rho = valid ? a.rho[ln] : 1.0;
  int iter = valid ? a.iter[ln] : 0;
  double prm[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) prm[k] = valid ? a.par[k * Bz + ln] : 0.0;

  // ---- load iterates: x into the C-fragment layout (registers), z / y into the tile buffers
  double x[2 * NT8], dsc[2 * NT8];
#pragma unroll
  for (int jt = 0; jt < NT8; ++jt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = 8 * jt + 2 * c + e;
      x[2 * jt + e] = (valid && col < N) ? a.xs[(size_t)ln * N + col] : 0.0;
      dsc[2 * jt + e] = (col < N) ? 1.0 / (1.0 + rho * lam[col]) : 0.0;
    }
  int flipf = 0;
  for (int r = 0; r < RM; ++r) {
    const int i = c + 4 * r;
    if (i < M) {
      zb[i] = valid ? a.zs[(size_t)ln * M + i] : 0.0;
      yb[i] = valid ? a.ys[(size_t)ln * M + i] : 0.0;
      db[i] = 0.0;
      if (pcode[i] == 2 && prm[4] * Ev[i] - lt[i] < MPCB_RHO_TOL) flipf = 1;
    }
  }
  if (valid && flipf) a.flip[ln] = 1;
  auto bounds = [&](int i, double &l_, double &u_) {            // this step's scaled bounds of row i
    l_ = lt[i];
    u_ = ut[i];
    const int pc = pcode[i];
    if (pc == 1) l_ = u_ = -((i == 0) ? prm[0] : (i == 1) ? prm[1] : (i == 2) ? prm[2] : prm[3]) * Ev[i];                     // rows 0..3: -x_hat
    else if (pc == 2) u_ = prm[4] * Ev[i];                      // velocity 1-norm bound
    else if (pc == 3) l_ = u_ = prm[5] * Ev[i];                 // disturbance pin
    else if (pc == 4) l_ = u_ = prm[6] * Ev[i];
  };
  auto rho_of = [&](int i) -> double {
    const uint8_t f = flags[i];
    return (f & 8) ? MPCB_RHO_MIN : ((f & 4) ? MPCB_RHO_EQ * rho : rho);
  };
  const double sigma = a.sigma, alpha = a.alpha, oma = 1.0 - a.alpha;

  // r base for the first iteration: sigma*x - q on the owned entries
#pragma unroll
  for (int jt = 0; jt < NT8; ++jt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = 8 * jt + 2 * c + e;
      if (col < LDN) nb[col] = (col < N) ? sigma * x[2 * jt + e] - qv[col] : 0.0;
    }
  __syncwarp();

  for (int it = 0; it < a.check_every; ++it) {
    // ---- P1: r += A'(rho_vec.*z - y) on columns j = c (mod 4)
    for (int r = 0; r < RN; ++r) {
      const int j = c + 4 * r;
      if (j < N) {
        double acc = 0.0;
        const int cn = ATn[j];
        for (int e = 0; e < cn; ++e) {
          const int i = ATc[e * NPAD + j];
          acc = fma(ATv[e * NPAD + j], rho_of(i) * zb[i] - yb[i], acc);
        }
        nb[j] += acc;
      }
    }
    __syncwarp();
    // ---- P2: w = dsc .* (V' r)
    double af[KS];
#pragma unroll
    for (int s = 0; s < KS; ++s) af[s] = nb[4 * s + c];
    __syncwarp();                                              // nb is free: every thread holds its A fragments
#pragma unroll
    for (int jt = 0; jt < NT8; ++jt) {
      double c0 = 0.0, c1 = 0.0;
      const double *vb = V + c * LDV + 8 * jt + g;             // B[k = 4s + c][col = 8jt + g]
#pragma unroll
      for (int s = 0; s < KS; ++s) dmma884(c0, c1, af[s], vb[4 * s * LDV]);
      const int col = 8 * jt + 2 * c;
      if (col < LDN) nb[col] = c0 * dsc[2 * jt];
      if (col + 1 < LDN) nb[col + 1] = c1 * dsc[2 * jt + 1];
    }
    __syncwarp();
    // ---- P3: x~ = V w ; x = alpha x~ + (1-alpha) x
#pragma unroll
    for (int s = 0; s < KS; ++s) af[s] = nb[4 * s + c];
    __syncwarp();
#pragma unroll
    for (int jt = 0; jt < NT8; ++jt) {
      double c0 = 0.0, c1 = 0.0;
      const double *vb = V + (8 * jt + g) * LDV + c;           // B[k = 4s + c][col = 8jt + g] = V[8jt + g][4s + c]
#pragma unroll
      for (int s = 0; s < KS; ++s) dmma884(c0, c1, af[s], vb[4 * s]);
      const int col = 8 * jt + 2 * c;
      if (col < LDN) nb[col] = c0;
      if (col + 1 < LDN) nb[col + 1] = c1;
      x[2 * jt] = alpha * c0 + oma * x[2 * jt];
      x[2 * jt + 1] = alpha * c1 + oma * x[2 * jt + 1];
    }
    __syncwarp();
    // ---- P4: z~ = A x~ on rows i = c (mod 4); projection; dual update
    for (int r = 0; r < RM; ++r) {
      const int i = c + 4 * r;
      if (i < M) {
        double zt = 0.0;
        const int an = An[i];
        for (int e = 0; e < an; ++e) zt = fma(Av[e * MPAD + i], nb[Ac[e * MPAD + i]], zt);
        double l_, u_;
        bounds(i, l_, u_);
        const double rvi = rho_of(i), zo = zb[i], yo = yb[i];
        const double zr = alpha * zt + oma * zo;
        const double zn = fmin(fmax(zr + (1.0 / rvi) * yo, l_), u_);
        const double dy = rvi * (zr - zn);
        yb[i] = yo + dy;
        zb[i] = zn;
        db[i] = dy;
      }
    }
    __syncwarp();
    // ---- r base of the next iteration (also the buffer the check reads x from when scaled by 1: see below)
#pragma unroll
    for (int jt = 0; jt < NT8; ++jt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = 8 * jt + 2 * c + e;
        if (col < LDN) nb[col] = (col < N) ? sigma * x[2 * jt + e] - qv[col] : 0.0;
      }
    __syncwarp();
  }
  iter += a.check_every;

  // =========================== update_info (OSQP auxil.c) ===========================
#pragma unroll
  for (int jt = 0; jt < NT8; ++jt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = 8 * jt + 2 * c + e;
      if (col < LDN) nb[col] = x[2 * jt + e];                   // cols >= N hold 0
    }
  __syncwarp();
  double pri_u = 0, nz_u = 0, nax_u = 0, pri_s = 0, nz_s = 0, nax_s = 0;
  for (int r = 0; r < RM; ++r) {
    const int i = c + 4 * r;
    if (i < M) {
      double Ax = 0.0;
      const int an = An[i];
      for (int e = 0; e < an; ++e) Ax = fma(Av[e * MPAD + i], nb[Ac[e * MPAD + i]], Ax);
      const double ei = Einv[i], zi = zb[i], pv = Ax - zi;
      pri_u = fmax(pri_u, fabs(ei * pv)); nz_u = fmax(nz_u, fabs(ei * zi)); nax_u = fmax(nax_u, fabs(ei * Ax));
      pri_s = fmax(pri_s, fabs(pv)); nz_s = fmax(nz_s, fabs(zi)); nax_s = fmax(nax_s, fabs(Ax));
    }
  }
  double dua_u = 0, npx_u = 0, naty_u = 0, dua_s = 0, npx_s = 0, naty_s = 0;
  for (int r = 0; r < RN; ++r) {
    const int j = c + 4 * r;
    if (j < N) {
      double Px = 0.0, Aty = 0.0;
      const int pn = Pn[j], cn = ATn[j];
      for (int e = 0; e < pn; ++e) Px = fma(Pv[e * NPAD + j], nb[Pc[e * NPAD + j]], Px);
      for (int e = 0; e < cn; ++e) Aty = fma(ATv[e * NPAD + j], yb[ATc[e * NPAD + j]], Aty);
      const double di = Dinv[j], dv = qv[j] + Px + Aty;
      dua_u = fmax(dua_u, fabs(di * dv)); npx_u = fmax(npx_u, fabs(di * Px)); naty_u = fmax(naty_u, fabs(di * Aty));
      dua_s = fmax(dua_s, fabs(dv)); npx_s = fmax(npx_s, fabs(Px)); naty_s = fmax(naty_s, fabs(Aty));
    }
  }
  pri_u = quad_max(pri_u); nz_u = quad_max(nz_u); nax_u = quad_max(nax_u);
  pri_s = quad_max(pri_s); nz_s = quad_max(nz_s); nax_s = quad_max(nax_s);
  dua_u = quad_max(dua_u) * a.cinv; npx_u = quad_max(npx_u); naty_u = quad_max(naty_u);
  dua_s = quad_max(dua_s); npx_s = quad_max(npx_s); naty_s = quad_max(naty_s);
  // primal-infeasibility certificate (is_primal_infeasible): project delta_y, A' delta_y
  double n1 = 0.0, l1 = 0.0;
  for (int r = 0; r < RM; ++r) {
    const int i = c + 4 * r;
    if (i < M) {
      const uint8_t f = flags[i];
      double d = db[i];
      if ((f & 3) == 3) d = 0.0;
      else if (f & 2) d = fmin(d, 0.0);
      else if (f & 1) d = fmax(d, 0.0);
      db[i] = d;
      double l_, u_;
      bounds(i, l_, u_);
      n1 = fmax(n1, fabs(Ev[i] * d));
      l1 += u_ * fmax(d, 0.0) + l_ * fmin(d, 0.0);
    }
  }
  __syncwarp();
  double n2 = 0.0;
  for (int r = 0; r < RN; ++r) {
    const int j = c + 4 * r;
    if (j < N) {
      double acc = 0.0;
      const int cn = ATn[j];
      for (int e = 0; e < cn; ++e) acc = fma(ATv[e * NPAD + j], db[ATc[e * NPAD + j]], acc);
      n2 = fmax(n2, fabs(Dinv[j] * acc));
    }
  }
  const double ndy = quad_max(n1), lhs = quad_sum(l1), natdy = quad_max(n2);

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
  int st = check(1.0);
  if (st == -10) {
    if (a.adaptive && (iter % a.adapt_interval == 0)) {
      const double pr = pri_s / (fmax(nz_s, nax_s) + 1e-10);
      const double du = dua_s / (fmax(a.qn_scaled, fmax(naty_s, npx_s)) + 1e-10);
      double est = rho * sqrt(pr / (du + 1e-10));
      est = fmin(fmax(est, MPCB_RHO_MIN), MPCB_RHO_MAX);
      if (est > rho * a.adapt_tol || est < rho / a.adapt_tol) rho = est;
    }
    if (iter >= a.max_iter) {
      st = check(10.0);
      if (st == -10) st = -2;
    }
  }

  // =========================== write back ===========================
  if (valid) {
#pragma unroll
    for (int jt = 0; jt < NT8; ++jt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = 8 * jt + 2 * c + e;
        if (col < N) a.xs[(size_t)ln * N + col] = x[2 * jt + e];
      }
    for (int r = 0; r < RM; ++r) {
      const int i = c + 4 * r;
      if (i < M) {
        a.zs[(size_t)ln * M + i] = zb[i];
        a.ys[(size_t)ln * M + i] = yb[i];
      }
    }
    if (c == 0) {
      a.rho[ln] = rho;
      a.iter[ln] = iter;
      a.status[ln] = st;
      if (st != -10) {
        a.u0[ln] = Dv[a.uoff] * nb[a.uoff];
        a.u0[Bz + ln] = Dv[a.uoff + 1] * nb[a.uoff + 1];
        a.lane_state[ln] = LANE_SOLVE_DONE;
      }
    }
  }
  const unsigned vm = __ballot_sync(0xffffffffu, valid && c == 0);
  if (lid == 0 && vm) atomicAdd(a.iter_total, (unsigned long long)__popc(vm) * (unsigned long long)a.check_every);
}
