// Generic per-lane path: trajectories whose constraint matrix changes CONTINUOUSLY from step to step,
// i.e. debris-avoidance lanes (reference src/simhelpers.py:48-64, 80-134: the half-plane row
// C[4,:] = [-slope, 1, 0, 0] of every LOS block is rebuilt from the estimate each control step).
//
// The reference then calls prob.update(Ax=...) (src/trajectorySimulate.py:348), on which OSQP re-runs
// its Ruiz equilibration from scratch and refactors the KKT matrix, so nothing can be shared across
// lanes or steps.  One CTA owns one lane and does, per control step, what OSQP does on the host:
//   geometry -> A values -> scale_data (10 Ruiz passes) -> bounds, constraint types, rho vector ->
//   M = P + sigma I + A' diag(rho_vec) A -> S = M^-1 (in-place Gauss-Jordan) -> ADMM with x~ = S r,
//   termination / infeasibility / adaptive rho as in team.cuh ->
//   controller select incl. the deadbeat avoidance law (:299-304), clip, plant, UKF, telemetry.
//
// The operator never leaves the SM.  Thread i owns ROW i of the n x n matrix (n <= 201): its first GEN_TMD = 128 entries
// live in TENSOR MEMORY (256 32-bit columns of the thread's own TMEM lane, shape 32x32b: a second register file, as in
// team.cuh -- there is no f64 tcgen05.mma), the rest in shared memory ([column][thread], conflict free).  Assembly, the
// Gauss-Jordan sweeps (pivot row broadcast through a double-buffered shared row; the owner of row k+1 publishes it while
// it applies pivot k, so a sweep costs two barriers) and the mat-vec all run on a thread's own row.  The first version of
// this kernel kept M in a global scratch matrix and spent ~90 % of a solve in L2 round trips of the sweeps
// (n^3 read-modify-writes per inverse): 4.6 k solves/s at Nx = 40.
// Parity with the oracle's scalar path: tests/test_gpu_parity.py::test_debris_*.
#pragma once
#include "common.cuh"
#include "sim.cuh"
#include "team.cuh"      // tensor-memory load / store wrappers

#define GEN_THREADS 256
// Optional cycle breakdown (-DGEN_PROFILE): thread 0 of every CTA accumulates clock64() deltas into tot[4..10]:
// 4 values + scaling, 5 assembly, 6 inversion, 7 iterations, 8 checks, 9 rest of the control step, 10 total.
#ifdef GEN_PROFILE
#define GP_DECL long long gp_t = clock64(); const long long gp_start = gp_t; long long gp_acc[6] = {0, 0, 0, 0, 0, 0};
#define GP_MARK(k) { const long long gp_n = clock64(); gp_acc[k] += gp_n - gp_t; gp_t = gp_n; }
#define GP_FLUSH if (tid == 0) { for (int q = 0; q < 6; ++q) atomicAdd(&a.tot[4 + q], (unsigned long long)gp_acc[q]); \
                                 atomicAdd(&a.tot[10], (unsigned long long)(clock64() - gp_start)); }
#else
#define GP_DECL
#define GP_MARK(k)
#define GP_FLUSH
#endif
#define GEN_TMD 128        // doubles of a row of S kept in tensor memory (2 x 128 = 256 columns; 8 warps x 256 = the SM's 512 x 4)

struct GenArgs {
  int n, m, nX, Nx, Nb, Nc, uoff, B, nnzA, nnzP, scaling;
  // sparsity pattern of A (CSR + CSC view), values for sign variant (+,+) and slope 1
  const int *rowptr, *colidx, *colptr, *rowidx, *cscpos;
  const double *baseA;        // [nnzA]
  const unsigned char *kindA; // 0 constant, 1 carries sign(vx_hat), 2 sign(vy_hat), 3 carries -slope
  int nlong, longcols[8];     // columns of A with more than 32 entries (the disturbance variables touch every dynamics row): a warp each in A'v
  const int *prow, *pcol;     // P in COO (both triangles), row-major
  const int *prowptr;         // [n+1] first COO entry of every row
  const double *pval, *q_u, *l_u, *u_u;
  double sigma, alpha, eps_abs, eps_rel, eps_pinf, adapt_tol, rho0;
  int check_every, adaptive, adapt_interval, max_iter;
  int mode, nsteps;
  // continuous simulator (trajectorySimulateC.py): RK4 substeps of T_cont, a solve every `ratio` substeps
  int ratio, n_sub_total, noise_hold_sub;
  double T_cont;
  SimConst sc;
  // debris geometry (src/mpcsim.py:99-123) and the deadbeat law (:190-203)
  int has_debris;
  double dcx, dcy, dside, ddetect, verts[8], Kd[8], Kid[2];
  SimOutDev out;
  const double *x0, *noise_in, *xhat;
  int n_refresh;
  double *xs, *zs, *ys, *rho, *u0;
  int *iter, *status;
  int warm;
  int *queue;
  unsigned long long *tot;
  double *stats;
};

struct GenLane {
  double xtrue[4], ux[6], uP[36], xstore[6], unext[2], noise[2], xfin[4], u0[2];
  double xintf, rho, slope, inter, val, dpin[2], xh[4];
  int step, iterm, succ, nsolve, ukf_clamp, fin, status, iter, lane, c1, c2, side_pos, bound_on;
  int sub;                   // continuous simulator: next substep index
};

// simhelpers.py:66-134 for one estimate `xe` (6, in the caller's storage order); fills the per-step QP data.
__device__ __forceinline__ void gen_geometry(const GenArgs &a, GenLane &L, const double *xe_in) {
  const SimConst &c = a.sc;
  double xe[6];
  for (int k = 0; k < 6; ++k) xe[k] = xe_in[k];
  L.c1 = (xe[2] >= 0) ? 1 : -1;
  L.c2 = (xe[3] >= 0) ? 1 : -1;
  for (int k = 0; k < 4; ++k) L.xh[k] = xe[k];          // the equality rows use the estimate BEFORE the in-track swap (:340-341)
  double xc0 = xe[0], xc1 = xe[1];                      // xestCalc
  double cx = a.dcx, cy = a.dcy;
  double x0s = xe[0], x1s = xe[1];                      // xest after the in-place swap
  if (c.in_track) {
    x0s = xe[1];
    x1s = xe[0];
    const double t = cx;
    cx = cy;
    cy = t;
  }
  const double hs = a.dside / 2;
  bool inside = false, near_ = false;
  double slope = 0.0, inter = 0.0;
  if (a.has_debris) {
    inside = (x0s - (cx + hs) < 0) && (x0s - (cx - hs) > 0);
    const int v = (x1s >= 0) ? (inside ? 1 : 0) : (inside ? 2 : 3);
    slope = (xc1 - a.verts[2 * v + 1]) / (xc0 - a.verts[2 * v]);
    inter = -slope * xc0 + xc1;
    near_ = (x0s - (cx + hs) < a.ddetect) && (x0s - (cx + hs) > 0);
  }
  L.slope = slope;
  L.inter = inter;
  L.val = fabs(xc0 - c.xr[0]) + fabs(xc1 - c.xr[1]);
  L.side_pos = (x1s >= 0) ? 1 : 0;
  L.bound_on = (inside || near_) ? 1 : 0;
  L.dpin[0] = c.is_reject ? xe[4] : 0.0;
  L.dpin[1] = c.is_reject ? xe[5] : 0.0;
}

__device__ __forceinline__ double gen_limit(double v) {
  if (v < 1e-4) return 1.0;
  return v > 1e4 ? 1e4 : v;
}

__global__ void __launch_bounds__(GEN_THREADS, 1) generic_lane_kernel(const __grid_constant__ GenArgs a) {
  extern __shared__ __align__(16) unsigned char gsm[];
  const int tid = threadIdx.x, lid = tid & 31, warp = tid >> 5;
  const int n = a.n, m = a.m, nnzA = a.nnzA, nnzP = a.nnzP;
  constexpr int T = GEN_THREADS, NW = GEN_THREADS / 32;
  // ---- shared memory: vectors of length n / m, scaled values, reductions, lane context
  double *p = reinterpret_cast<double *>(gsm);
  auto take = [&](int cnt) { double *r = p; p += (cnt + 1) & ~1; return r; };
  double *x = take(n), *xt = take(n), *rb = take(n + 4), *qs = take(n), *D = take(n), *Dinv = take(n), *cmax = take(n);
  double *z = take(m), *y = take(m), *v = take(m), *dyb = take(m), *lo = take(m), *hi = take(m), *rv = take(m), *E = take(m),
         *Einv = take(m), *rmax = take(m);
  double *As = take(nnzA), *Ps = take(nnzP), *red = take(16 * NW), *fcol = take(n);
  int *ctype = reinterpret_cast<int *>(take((m + 1) / 2 + 1));
  // operator storage: row tid of S = tmd doubles in tensor memory + (npad - tmd) in Sx[.][tid]; rowb = two pivot-row buffers
  const int npad = (n + 3) & ~3, tmd = npad < GEN_TMD ? npad : GEN_TMD, TS = (n + 31) & ~31;
  double *rowb = take(2 * npad), *Sx = take((npad - tmd) * TS);
  GenLane &L = *reinterpret_cast<GenLane *>(p);
  // sparsity pattern of A in shared memory (uint16: m, n, nnz < 65536): the per-iteration products walk these lists, and from
  // global memory every walk is a chain of dependent L1 / L2 round trips
  uint16_t *ixp = reinterpret_cast<uint16_t *>(reinterpret_cast<unsigned char *>(p) + ((sizeof(GenLane) + 15) & ~(size_t)15));
  uint16_t *ix_rowptr = ixp, *ix_colptr = ix_rowptr + ((m + 2) & ~1), *ix_colidx = ix_colptr + ((n + 2) & ~1),
           *ix_rowidx = ix_colidx + ((nnzA + 1) & ~1), *ix_cscpos = ix_rowidx + ((nnzA + 1) & ~1);
  for (int i = tid; i <= m; i += T) ix_rowptr[i] = (uint16_t)a.rowptr[i];
  for (int j = tid; j <= n; j += T) ix_colptr[j] = (uint16_t)a.colptr[j];
  for (int e = tid; e < nnzA; e += T) {
    ix_colidx[e] = (uint16_t)a.colidx[e];
    ix_rowidx[e] = (uint16_t)a.rowidx[e];
    ix_cscpos[e] = (uint16_t)a.cscpos[e];
  }
  __shared__ int s_lane;
  __shared__ double s_c;
  __shared__ uint32_t s_tmem;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tmem)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = s_tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(2 * GEN_TMD * (warp >> 2));
  if (tid < 4) rb[n + tid] = 0.0;        // the mat-vec reads r in chunks of four: the padding multiplies zero columns of S
  const bool own = tid < n;              // owns a row of S (every thread still executes the warp-collective TMEM ops)
  // entry j of the thread's own row (j uniform across the warp)
  auto row_get = [&](int j) -> double {
    if (j < tmd) {
      uint32_t c4[4];
      tmem_ld4(taddr + 2 * (j & ~1), c4);
      tmem_wait_ld4(c4);
      return (j & 1) ? u2d(c4[2], c4[3]) : u2d(c4[0], c4[1]);
    }
    return own ? Sx[(size_t)(j - tmd) * TS + tid] : 0.0;
  };

  auto team_max = [&](double val) -> double {
    val = warp_max_nonneg(val);
    __syncthreads();
    if (lid == 0) red[warp] = val;
    __syncthreads();
    double r = red[0];
    for (int w = 1; w < NW; ++w) r = fmax(r, red[w]);
    __syncthreads();
    return r;
  };
  auto team_sum = [&](double val) -> double {
    val = warp_sum(val);
    __syncthreads();
    if (lid == 0) red[warp] = val;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < NW; ++w) r += red[w];
    __syncthreads();
    return r;
  };
  auto Arow = [&](int i, const double *vec) -> double {        // (A_s vec)[i]
    double acc = 0.0;
    for (int e = ix_rowptr[i]; e < ix_rowptr[i + 1]; ++e) acc = fma(As[e], vec[ix_colidx[e]], acc);
    return acc;
  };
  auto ATcol = [&](int j, const double *vec) -> double {       // (A_s' vec)[j]
    double acc = 0.0;
    for (int e = ix_colptr[j]; e < ix_colptr[j + 1]; ++e) acc = fma(As[ix_cscpos[e]], vec[ix_rowidx[e]], acc);
    return acc;
  };

  unsigned long long my_iters = 0, my_rebuilds = 0;
  GP_DECL
  while (true) {
    __syncthreads();
    if (tid == 0) s_lane = atomicAdd(a.queue, 1);
    __syncthreads();
    const int ln = s_lane;
    if (ln >= a.B) break;
    const size_t B = a.B;
    // ---------------- lane initial conditions (trajectorySimulate.py:248-269)
    if (tid == 0) {
      L.lane = ln; L.rho = a.rho0; L.xintf = 0.0; L.step = 0; L.succ = 0; L.nsolve = 0; L.ukf_clamp = 0; L.fin = 0;
      L.status = -10; L.iter = 0; L.u0[0] = L.u0[1] = 0.0;
      double xe[6];
      if (a.mode == MODE_QP_ONLY) {
        for (int k = 0; k < 6; ++k) xe[k] = a.xhat[k * B + ln];
        if (a.warm) L.rho = a.rho[ln];
        L.iterm = 0;
      } else {
        for (int k = 0; k < 4; ++k) xe[k] = a.x0[k * B + ln];
        xe[4] = xe[5] = 0.0;
        for (int k = 0; k < 4; ++k) { L.xtrue[k] = xe[k]; L.xfin[k] = nan(""); }
        for (int k = 0; k < 6; ++k) { L.ux[k] = xe[k]; L.xstore[k] = xe[k]; }
        for (int k = 0; k < 36; ++k) L.uP[k] = (k % 7 == 0) ? ((k / 7 < 4) ? 1e-20 : 1.0) : 0.0;
        L.unext[0] = L.unext[1] = 0.0;
        L.noise[0] = (a.sc.has_noise && a.noise_in) ? a.noise_in[ln] : 0.0;
        L.noise[1] = (a.sc.has_noise && a.noise_in) ? a.noise_in[B + ln] : 0.0;
        const int T1 = a.out.T1;
        if (a.out.x_true) for (int k = 0; k < 4; ++k) a.out.x_true[((size_t)k * T1) * B + ln] = xe[k];
        if (a.out.x_est) for (int k = 0; k < 6; ++k) a.out.x_est[((size_t)k * T1) * B + ln] = xe[k];
        if (a.out.ctrl) for (int k = 0; k < 2; ++k) a.out.ctrl[((size_t)k * T1) * B + ln] = 0.0;
        if (a.mode == MODE_CONTINUOUS) {     // same prologue as init_kernel (sim.cuh): ratio substeps of x0 with zero control
          const size_t NS = a.out.NS;
          const int nfill = min(a.ratio + 1, a.n_sub_total);
          for (int sdx = 0; sdx < nfill; ++sdx) {
            if (a.out.x_true_sub) for (int k = 0; k < 4; ++k) a.out.x_true_sub[((size_t)k * NS + sdx) * B + ln] = xe[k];
            if (a.out.ctrl_sub) for (int k = 0; k < 2; ++k) a.out.ctrl_sub[((size_t)k * NS + sdx) * B + ln] = 0.0;
          }
          L.sub = a.ratio;
          L.iterm = a.n_sub_total;
          if (a.ratio >= a.n_sub_total - 1) L.fin = 1;
          else if (terminated(a.sc, xe)) { L.iterm = a.ratio; L.fin = 1; }
        } else {
          L.iterm = a.nsteps;
          if (a.nsteps <= 0) L.fin = 1;
          else if (terminated(a.sc, xe)) { L.iterm = 0; L.fin = 1; }
        }
      }
      gen_geometry(a, L, xe);
    }
    for (int j = tid; j < n; j += T) x[j] = a.warm ? a.xs[(size_t)ln * n + j] : 0.0;
    for (int i = tid; i < m; i += T) {
      z[i] = a.warm ? a.zs[(size_t)ln * m + i] : 0.0;
      y[i] = a.warm ? a.ys[(size_t)ln * m + i] : 0.0;
    }
    __syncthreads();

    while (!L.fin) {
      // ================= prob.update(l,u); prob.update(Ax,l,u): values, scale_data, bounds, types =================
      double rho = L.rho;
      for (int e = tid; e < nnzA; e += T) {
        const int k = a.kindA[e];
        As[e] = a.baseA[e] * (k == 1 ? (double)L.c1 : k == 2 ? (double)L.c2 : k == 3 ? L.slope : 1.0);
      }
      for (int e = tid; e < nnzP; e += T) Ps[e] = a.pval[e];
      for (int j = tid; j < n; j += T) { qs[j] = a.q_u[j]; D[j] = 1.0; }
      for (int i = tid; i < m; i += T) E[i] = 1.0;
      if (tid == 0) s_c = 1.0;
      __syncthreads();
      for (int pass = 0; pass < a.scaling; ++pass) {              // OSQP scale_data (oracle/osqp_ref.py ruiz_scale)
        for (int j = tid; j < n; j += T) cmax[j] = 0.0;
        for (int i = tid; i < m; i += T) rmax[i] = 0.0;
        __syncthreads();
        for (int e = tid; e < nnzP; e += T)
          atomicMax(reinterpret_cast<unsigned long long *>(&cmax[a.pcol[e]]), (unsigned long long)__double_as_longlong(fabs(Ps[e])));
        for (int i = tid; i < m; i += T) {
          double rm = 0.0;
          for (int e = ix_rowptr[i]; e < ix_rowptr[i + 1]; ++e) {
            const double av = fabs(As[e]);
            rm = fmax(rm, av);
            atomicMax(reinterpret_cast<unsigned long long *>(&cmax[ix_colidx[e]]), (unsigned long long)__double_as_longlong(av));
          }
          rmax[i] = rm;
        }
        __syncthreads();
        for (int j = tid; j < n; j += T) cmax[j] = 1.0 / sqrt(gen_limit(cmax[j]));
        for (int i = tid; i < m; i += T) rmax[i] = 1.0 / sqrt(gen_limit(rmax[i]));
        __syncthreads();
        for (int e = tid; e < nnzP; e += T) Ps[e] = cmax[a.prow[e]] * Ps[e] * cmax[a.pcol[e]];
        for (int i = tid; i < m; i += T) {
          for (int e = ix_rowptr[i]; e < ix_rowptr[i + 1]; ++e) As[e] = rmax[i] * As[e] * cmax[ix_colidx[e]];
          E[i] *= rmax[i];
        }
        for (int j = tid; j < n; j += T) { qs[j] *= cmax[j]; D[j] *= cmax[j]; }
        __syncthreads();
        // cost normalisation: c_tmp = 1 / limit(max(mean_j ||P_:j||inf, limit(||q||inf)))
        for (int j = tid; j < n; j += T) cmax[j] = 0.0;
        __syncthreads();
        for (int e = tid; e < nnzP; e += T)
          atomicMax(reinterpret_cast<unsigned long long *>(&cmax[a.pcol[e]]), (unsigned long long)__double_as_longlong(fabs(Ps[e])));
        __syncthreads();
        double part = 0.0, qn = 0.0;
        for (int j = tid; j < n; j += T) { part += cmax[j]; qn = fmax(qn, fabs(qs[j])); }
        const double csum = team_sum(part);
        qn = team_max(qn);
        double ct = fmax(csum / n, gen_limit(qn));
        ct = 1.0 / gen_limit(ct);
        for (int e = tid; e < nnzP; e += T) Ps[e] *= ct;
        for (int j = tid; j < n; j += T) qs[j] *= ct;
        if (tid == 0) s_c *= ct;
        __syncthreads();
      }
      const double cinv = 1.0 / s_c;
      double qn_u = 0.0, qn_s = 0.0;
      for (int j = tid; j < n; j += T) {
        Dinv[j] = 1.0 / D[j];
        qn_u = fmax(qn_u, fabs(Dinv[j] * qs[j]));
        qn_s = fmax(qn_s, fabs(qs[j]));
      }
      qn_u = team_max(qn_u);
      qn_s = team_max(qn_s);
      for (int i = tid; i < m; i += T) {
        Einv[i] = 1.0 / E[i];
        double l_ = a.l_u[i], u_ = a.u_u[i];
        if (i < 4) l_ = u_ = -L.xh[i];
        else if (i >= m - 2) l_ = u_ = L.dpin[i - (m - 2)];
        else if (i >= a.nX && i < a.nX + 5 * (a.Nb + 1)) {
          const int jj = (i - a.nX) % 5;
          if (jj == 3) u_ = L.val;
          else if (jj == 4 && L.bound_on) {
            if (L.side_pos) l_ = L.inter; else u_ = L.inter;
          }
        }
        l_ = E[i] * fmax(l_, -1e30);
        u_ = E[i] * fmin(u_, 1e30);
        lo[i] = l_;
        hi[i] = u_;
        const bool fr = (l_ < -1e26) && (u_ > 1e26);
        ctype[i] = fr ? -1 : ((u_ - l_ < MPCB_RHO_TOL) ? 1 : 0);
      }
      __syncthreads();
      GP_MARK(0)
      int iter = 0, st = -10;
      bool need_op = true;                                        // update(Ax) always refactors

      while (st == -10) {
        if (need_op) {
          for (int i = tid; i < m; i += T) rv[i] = ctype[i] == -1 ? MPCB_RHO_MIN : (ctype[i] == 1 ? MPCB_RHO_EQ * rho : rho);
          // M = P + sigma I + A' diag(rho_vec) A = sum over rows r of rho_vec_r a_r a_r'.  Thread i builds row i of M in a
          // thread-local array: for every row r of A that touches column i (CSC list, ascending r: a FIXED summation order)
          // it adds rho_vec_r A_ri A_rj for the entries j of that row, then the row goes to tensor / shared memory four
          // entries at a time.  (Merging column i with every column j instead costs the two disturbance columns, which
          // touch all 4 (Nx + 1) dynamics rows, 35 k steps per assembly.)
          __syncthreads();
          {
            double Mrow[GEN_THREADS];      // n <= GEN_THREADS; local memory (L1 / L2), dynamic index
            for (int j = 0; j < npad; ++j) Mrow[j] = 0.0;
            if (own) {
              Mrow[tid] = a.sigma;
              for (int ea = ix_colptr[tid]; ea < ix_colptr[tid + 1]; ++ea) {
                const int r = ix_rowidx[ea];
                const double w = rv[r] * As[ix_cscpos[ea]];
                for (int e = ix_rowptr[r]; e < ix_rowptr[r + 1]; ++e) {
                  const int j = ix_colidx[e];
                  Mrow[j] = fma(w, As[e], Mrow[j]);
                }
              }
              for (int pe = a.prowptr[tid]; pe < a.prowptr[tid + 1]; ++pe) Mrow[a.pcol[pe]] += Ps[pe];
            }
            for (int j0 = 0; j0 < npad; j0 += 4) {
              if (j0 < tmd) {
                const uint32_t w8[8] = {(uint32_t)__double2loint(Mrow[j0]), (uint32_t)__double2hiint(Mrow[j0]),
                                        (uint32_t)__double2loint(Mrow[j0 + 1]), (uint32_t)__double2hiint(Mrow[j0 + 1]),
                                        (uint32_t)__double2loint(Mrow[j0 + 2]), (uint32_t)__double2hiint(Mrow[j0 + 2]),
                                        (uint32_t)__double2loint(Mrow[j0 + 3]), (uint32_t)__double2hiint(Mrow[j0 + 3])};
                tmem_st8(taddr + 2 * j0, w8);
              } else if (own) {
#pragma unroll
                for (int q = 0; q < 4; ++q) Sx[(size_t)(j0 + q - tmd) * TS + tid] = Mrow[j0 + q];
              }
            }
            tmem_wait_st();
          }
          GP_MARK(1)
          // in-place Gauss-Jordan inversion (M is symmetric positive definite: no pivoting).  rowb[k & 1] holds row k as it
          // stands before sweep k (unscaled): row 0 is published here, row k+1 by its owner during sweep k.
          for (int j0 = 0; j0 < npad; j0 += 4) {
            double v4[4];
            if (j0 < tmd) {
              uint32_t c8[8];
              tmem_ld8(taddr + 2 * j0, c8);
              tmem_wait_ld8(c8);
#pragma unroll
              for (int q = 0; q < 4; ++q) v4[q] = u2d(c8[2 * q], c8[2 * q + 1]);
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) v4[q] = own ? Sx[(size_t)(j0 + q - tmd) * TS + tid] : 0.0;
            }
            if (tid == 0) {
#pragma unroll
              for (int q = 0; q < 4; ++q) rowb[j0 + q] = v4[q];
            }
          }
          for (int k = 0; k < n; ++k) {
            double *buf = rowb + (k & 1) * npad, *nxt = rowb + ((k + 1) & 1) * npad;
            __syncthreads();
            // every thread scales on the fly (no separate pass over the pivot row, one barrier per sweep):
            //   row k   <- S_kj / S_kk, S_kk <- 1 / S_kk;      row i <- S_ij - (S_ik / S_kk) S_kj, S_ik <- -S_ik / S_kk
            const double piv = 1.0 / buf[k];
            const double g = row_get(k) * piv;                   // S_ik / S_kk
            const bool is_k = tid == k, is_next = tid == k + 1;
            const double mul = is_k ? piv : -g, diag = is_k ? piv : -g, keep = is_k ? 0.0 : 1.0;
            // One DMUL + one DFMA per entry for owner and non-owner alike; only the group that holds the pivot column pays
            // the compare / select (ncu: FSEL + ISETP used to dominate this loop, 68 % of the kernel's samples).
            auto updf = [&](double b, double old) -> double { return fma(mul, b, keep * old); };
            auto upds = [&](int j, double b, double old) -> double { return (j == k) ? diag : fma(mul, b, keep * old); };
            const double2 *buf2 = reinterpret_cast<const double2 *>(buf);
            double2 *nxt2 = reinterpret_cast<double2 *>(nxt);
            int j0 = 0;
            // tensor-memory part, 32 entries per step: four 16-column loads in flight, one wait, four stores
            for (; j0 + 32 <= tmd; j0 += 32) {
              uint32_t q[4][16];
              tmem_ld16(taddr + 2 * j0, q[0]);
              tmem_ld16(taddr + 2 * j0 + 16, q[1]);
              tmem_ld16(taddr + 2 * j0 + 32, q[2]);
              tmem_ld16(taddr + 2 * j0 + 48, q[3]);
              double2 b2[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) b2[i] = buf2[j0 / 2 + i];
              tmem_wait_ld2(q[0], q[1]);
              tmem_wait_ld2(q[2], q[3]);
              double2 r2[16];
              if ((unsigned)(k - j0) < 32u) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const int gq = i >> 2, e = (i & 3) * 4;        // entries 2i, 2i+1 of the group = registers e..e+3 of load gq
                  r2[i].x = upds(j0 + 2 * i, b2[i].x, u2d(q[gq][e], q[gq][e + 1]));
                  r2[i].y = upds(j0 + 2 * i + 1, b2[i].y, u2d(q[gq][e + 2], q[gq][e + 3]));
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const int gq = i >> 2, e = (i & 3) * 4;
                  r2[i].x = updf(b2[i].x, u2d(q[gq][e], q[gq][e + 1]));
                  r2[i].y = updf(b2[i].y, u2d(q[gq][e + 2], q[gq][e + 3]));
                }
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int gq = i >> 2, e = (i & 3) * 4;
                q[gq][e] = (uint32_t)__double2loint(r2[i].x); q[gq][e + 1] = (uint32_t)__double2hiint(r2[i].x);
                q[gq][e + 2] = (uint32_t)__double2loint(r2[i].y); q[gq][e + 3] = (uint32_t)__double2hiint(r2[i].y);
              }
              tmem_st16(taddr + 2 * j0, q[0]);
              tmem_st16(taddr + 2 * j0 + 16, q[1]);
              tmem_st16(taddr + 2 * j0 + 32, q[2]);
              tmem_st16(taddr + 2 * j0 + 48, q[3]);
              if (is_next) {
#pragma unroll
                for (int i = 0; i < 16; ++i) nxt2[j0 / 2 + i] = r2[i];
              }
            }
            for (; j0 < tmd; j0 += 4) {                          // tail of the tensor-memory part
              uint32_t c8[8];
              tmem_ld8(taddr + 2 * j0, c8);
              const double2 ba = buf2[j0 / 2], bb = buf2[j0 / 2 + 1];
              tmem_wait_ld8(c8);
              double2 ra, rb2;
              ra.x = upds(j0, ba.x, u2d(c8[0], c8[1]));
              ra.y = upds(j0 + 1, ba.y, u2d(c8[2], c8[3]));
              rb2.x = upds(j0 + 2, bb.x, u2d(c8[4], c8[5]));
              rb2.y = upds(j0 + 3, bb.y, u2d(c8[6], c8[7]));
              c8[0] = (uint32_t)__double2loint(ra.x); c8[1] = (uint32_t)__double2hiint(ra.x);
              c8[2] = (uint32_t)__double2loint(ra.y); c8[3] = (uint32_t)__double2hiint(ra.y);
              c8[4] = (uint32_t)__double2loint(rb2.x); c8[5] = (uint32_t)__double2hiint(rb2.x);
              c8[6] = (uint32_t)__double2loint(rb2.y); c8[7] = (uint32_t)__double2hiint(rb2.y);
              tmem_st8(taddr + 2 * j0, c8);
              if (is_next) { nxt2[j0 / 2] = ra; nxt2[j0 / 2 + 1] = rb2; }
            }
            if (own) {                                           // shared-memory part, four entries per step
              // (pointer-stepped and unrolled: ncu put 38 % of the kernel's warp instructions in this loop's index arithmetic)
              double *sx = Sx + tid;
              const double2 *bp = buf2 + tmd / 2;
              double2 *np2 = nxt2 + tmd / 2;
              const int nq = (npad - tmd) >> 2, kq = (k >= tmd) ? ((k - tmd) >> 2) : -1;
              const size_t st4 = (size_t)4 * TS;
#pragma unroll 2
              for (int q = 0; q < nq; ++q, sx += st4, bp += 2, np2 += 2) {
                const double2 ba = bp[0], bb = bp[1];
                const double o0 = sx[0], o1 = sx[TS], o2 = sx[2 * TS], o3 = sx[3 * TS];
                double2 ra, rb2;
                if (q == kq) {
                  const int jb = tmd + 4 * q;
                  ra.x = upds(jb, ba.x, o0);
                  ra.y = upds(jb + 1, ba.y, o1);
                  rb2.x = upds(jb + 2, bb.x, o2);
                  rb2.y = upds(jb + 3, bb.y, o3);
                } else {
                  ra.x = updf(ba.x, o0);
                  ra.y = updf(ba.y, o1);
                  rb2.x = updf(bb.x, o2);
                  rb2.y = updf(bb.y, o3);
                }
                sx[0] = ra.x; sx[TS] = ra.y; sx[2 * TS] = rb2.x; sx[3 * TS] = rb2.y;
                if (is_next) { np2[0] = ra; np2[1] = rb2; }
              }
            }
            tmem_wait_st();
          }
          __syncthreads();
          need_op = false;
          ++my_rebuilds;
          GP_MARK(2)
        }
        for (int i = tid; i < m; i += T) v[i] = rv[i] * z[i] - y[i];
        __syncthreads();
        for (int it = 0; it < a.check_every; ++it) {
          for (int j = tid; j < n; j += T)
            if (ix_colptr[j + 1] - ix_colptr[j] <= 32) rb[j] = a.sigma * x[j] - qs[j] + ATcol(j, v);
          for (int q = warp; q < a.nlong; q += NW) {              // long columns: 32 lanes share the entries, fixed butterfly sum
            const int j = a.longcols[q];
            double acc = 0.0;
            for (int e = ix_colptr[j] + lid; e < ix_colptr[j + 1]; e += 32) acc = fma(As[ix_cscpos[e]], v[ix_rowidx[e]], acc);
            acc = warp_sum(acc);
            if (lid == 0) rb[j] = a.sigma * x[j] - qs[j] + acc;
          }
          __syncthreads();
          {                                  // x~ = S r: a thread's own row, tensor-memory part two chunks in flight
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int j0 = 0;
            if (tmd >= 32) {             // 32 entries per step in two halves: the next half's loads fly while this half's FMAs issue
              uint32_t A0[16], A1[16], B0[16], B1[16];
              auto use16 = [&](const uint32_t (&c0)[16], const uint32_t (&c1)[16], int jb) {
                const double2 *r2 = reinterpret_cast<const double2 *>(rb + jb);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const double2 ra = r2[i], rc = r2[4 + i];
                  a0 = fma(u2d(c0[4 * i], c0[4 * i + 1]), ra.x, a0);
                  a1 = fma(u2d(c0[4 * i + 2], c0[4 * i + 3]), ra.y, a1);
                  a2 = fma(u2d(c1[4 * i], c1[4 * i + 1]), rc.x, a2);
                  a3 = fma(u2d(c1[4 * i + 2], c1[4 * i + 3]), rc.y, a3);
                }
              };
              tmem_ld16(taddr, A0); tmem_ld16(taddr + 16, A1);
              tmem_ld16(taddr + 32, B0); tmem_ld16(taddr + 48, B1);
              tmem_wait_ld2(A0, A1);
              tmem_wait_ld2(B0, B1);
              for (; j0 + 32 <= tmd; j0 += 32) {
                use16(A0, A1, j0);
                const bool more = j0 + 64 <= tmd;
                if (more) { tmem_ld16(taddr + 2 * (j0 + 32), A0); tmem_ld16(taddr + 2 * (j0 + 32) + 16, A1); }
                use16(B0, B1, j0 + 16);
                if (more) { tmem_ld16(taddr + 2 * (j0 + 48), B0); tmem_ld16(taddr + 2 * (j0 + 48) + 16, B1); }
                tmem_wait_ld2(A0, A1);
                tmem_wait_ld2(B0, B1);
              }
            }
            for (; j0 < tmd; j0 += 8) {  // tail: eight (or the last four) entries per step
              uint32_t ca[8], cb[8];
              tmem_ld8(taddr + 2 * j0, ca);
              const bool two = j0 + 4 < tmd;
              if (two) tmem_ld8(taddr + 2 * j0 + 8, cb);
              else {
#pragma unroll
                for (int q = 0; q < 8; ++q) cb[q] = 0u;
              }
              tmem_wait_ld8x2(ca, cb);
              const double2 r01 = *reinterpret_cast<const double2 *>(rb + j0), r23 = *reinterpret_cast<const double2 *>(rb + j0 + 2);
              a0 = fma(u2d(ca[0], ca[1]), r01.x, a0);
              a1 = fma(u2d(ca[2], ca[3]), r01.y, a1);
              a2 = fma(u2d(ca[4], ca[5]), r23.x, a2);
              a3 = fma(u2d(ca[6], ca[7]), r23.y, a3);
              if (two) {
                const double2 r45 = *reinterpret_cast<const double2 *>(rb + j0 + 4), r67 = *reinterpret_cast<const double2 *>(rb + j0 + 6);
                a0 = fma(u2d(cb[0], cb[1]), r45.x, a0);
                a1 = fma(u2d(cb[2], cb[3]), r45.y, a1);
                a2 = fma(u2d(cb[4], cb[5]), r67.x, a2);
                a3 = fma(u2d(cb[6], cb[7]), r67.y, a3);
              }
            }
            if (own) {
              for (int j = tmd; j < npad; j += 4) {
                const double *sx = Sx + (size_t)(j - tmd) * TS + tid;
                a0 = fma(sx[0], rb[j], a0);
                a1 = fma(sx[TS], rb[j + 1], a1);
                a2 = fma(sx[2 * TS], rb[j + 2], a2);
                a3 = fma(sx[3 * TS], rb[j + 3], a3);
              }
              xt[tid] = (a0 + a1) + (a2 + a3);
            }
          }
          __syncthreads();
          for (int j = tid; j < n; j += T) x[j] = a.alpha * xt[j] + (1.0 - a.alpha) * x[j];
          for (int i = tid; i < m; i += T) {
            const double zt = Arow(i, xt);
            const double zr = a.alpha * zt + (1.0 - a.alpha) * z[i];
            const double zn = fmin(fmax(zr + (1.0 / rv[i]) * y[i], lo[i]), hi[i]);
            const double dy = rv[i] * (zr - zn);
            y[i] += dy;
            z[i] = zn;
            v[i] = rv[i] * zn - y[i];
            dyb[i] = dy;
          }
          __syncthreads();
        }
        iter += a.check_every;
        my_iters += (unsigned long long)a.check_every;
        GP_MARK(3)
        // ---- update_info / check_termination / adapt_rho (same arithmetic as team.cuh)
        for (int i = tid; i < m; i += T) v[i] = y[i];
        __syncthreads();
        double mu[12];
        for (int q = 0; q < 12; ++q) mu[q] = 0.0;
        double l1 = 0.0;
        for (int i = tid; i < m; i += T) {
          const double Ax = Arow(i, x), pv = Ax - z[i], ei = Einv[i];
          mu[0] = fmax(mu[0], fabs(ei * pv)); mu[1] = fmax(mu[1], fabs(ei * z[i])); mu[2] = fmax(mu[2], fabs(ei * Ax));
          mu[3] = fmax(mu[3], fabs(pv)); mu[4] = fmax(mu[4], fabs(z[i])); mu[5] = fmax(mu[5], fabs(Ax));
        }
        for (int j = tid; j < n; j += T) {                                                       // xt <- P x, fixed order
          double acc = 0.0;
          for (int e = a.prowptr[j]; e < a.prowptr[j + 1]; ++e) acc = fma(Ps[e], x[a.pcol[e]], acc);
          xt[j] = acc;
        }
        __syncthreads();
        for (int j = tid; j < n; j += T) {
          const double Px = xt[j], Aty = ATcol(j, v), dv = qs[j] + Px + Aty, di = Dinv[j];
          mu[6] = fmax(mu[6], fabs(di * dv)); mu[7] = fmax(mu[7], fabs(di * Px)); mu[8] = fmax(mu[8], fabs(di * Aty));
          mu[9] = fmax(mu[9], fabs(dv)); mu[10] = fmax(mu[10], fabs(Px)); mu[11] = fmax(mu[11], fabs(Aty));
        }
        for (int q = 0; q < 12; ++q) mu[q] = team_max(mu[q]);
        // primal-infeasibility certificate
        double n1 = 0.0;
        for (int i = tid; i < m; i += T) {
          double d = dyb[i];
          const bool il = lo[i] < -1e26, iu = hi[i] > 1e26;
          if (il && iu) d = 0.0;
          else if (iu) d = fmin(d, 0.0);
          else if (il) d = fmax(d, 0.0);
          v[i] = d;
          n1 = fmax(n1, fabs(E[i] * d));
          l1 += hi[i] * fmax(d, 0.0) + lo[i] * fmin(d, 0.0);
        }
        const double ndy = team_max(n1);
        const double lhs = team_sum(l1);
        double n2 = 0.0;
        for (int j = tid; j < n; j += T) n2 = fmax(n2, fabs(Dinv[j] * ATcol(j, v)));
        const double natdy = team_max(n2);
        auto check = [&](double k) -> int {
          const double eps_p = k * a.eps_abs + k * a.eps_rel * fmax(mu[1], mu[2]);
          const double eps_d = k * a.eps_abs + k * a.eps_rel * cinv * fmax(qn_u, fmax(mu[8], mu[7]));
          const bool prim_ok = mu[0] < eps_p, dual_ok = mu[6] * cinv < eps_d;
          if (prim_ok && dual_ok) return (k > 1.0) ? 2 : 1;
          if (!prim_ok) {
            const double eps_i = k * a.eps_pinf;
            if (ndy > MPCB_DIV_TOL && lhs < -eps_i * ndy && natdy < eps_i * ndy) return (k > 1.0) ? 3 : -3;
          }
          return -10;
        };
        st = check(1.0);
        if (st == -10) {
          if (a.adaptive && (iter % a.adapt_interval == 0)) {
            const double pr = mu[3] / (fmax(mu[4], mu[5]) + 1e-10);
            const double du = mu[9] / (fmax(qn_s, fmax(mu[11], mu[10])) + 1e-10);
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
        __syncthreads();
      }

      GP_MARK(4)
      // ================= rest of the control step (thread 0; trajectorySimulate.py:298-356) =================
      if (tid == 0) {
        const SimConst &c = a.sc;
        L.rho = rho; L.status = st; L.iter = iter;
        L.u0[0] = D[a.uoff] * x[a.uoff];
        L.u0[1] = D[a.uoff + 1] * x[a.uoff + 1];
        if (a.mode == MODE_QP_ONLY) {
          L.fin = 1;
        } else {
          const int T1 = a.out.T1, i = L.step;
          const double *xs4 = L.xstore;                          // xestO[:4, i] (x/y swapped for in-track)
          double u[2], uraw[2], xn[4];
          int code;
          const double hs = a.dside / 2;
          if (st != 1) {
            const bool in_box = a.has_debris && (xs4[0] - (a.dcx + hs) < 0) && (xs4[0] - (a.dcx - hs) > 0) &&
                                (xs4[1] < a.dcy + hs) && (xs4[1] > a.dcy - hs);
            if (in_box) {                                        // deadbeat collision avoidance (:300-304)
              L.xintf = L.xintf + xs4[1] - (a.dcy + hs);
              for (int r = 0; r < 2; ++r) {
                double acc = 0.0;
                for (int j = 0; j < 4; ++j) acc += a.Kd[r * 4 + j] * xs4[j];
                u[r] = -acc - a.Kid[r] * L.xintf;
              }
              code = 3;
            } else {                                             // failsafe (homing) LQR (:305-309)
              L.xintf = L.xintf + xs4[0] - c.xr[0];
              for (int r = 0; r < 2; ++r) {
                double acc = 0.0;
                for (int j = 0; j < 4; ++j) acc += c.Kpf[r * 4 + j] * xs4[j];
                u[r] = -acc - c.Kif[r] * L.xintf;
              }
              code = 2;
            }
          } else {
            L.xintf = 0.0;
            u[0] = L.u0[0];
            u[1] = L.u0[1];
            code = 1;
          }
          uraw[0] = u[0];
          uraw[1] = u[1];
          const double nrm = sqrt(u[0] * u[0] + u[1] * u[1]);
          if (nrm > c.umax0) {
            u[0] = u[0] * (c.umax0 / nrm);
            const double nrm2 = sqrt(u[0] * u[0] + u[1] * u[1]);
            u[1] = u[1] * (c.umax0 / nrm2);
          }
          const double uprev[2] = {L.unext[0], L.unext[1]};
          L.nsolve += 1;
          if (a.out.status) a.out.status[(size_t)i * B + ln] = (int8_t)st;
          if (a.out.iters) a.out.iters[(size_t)i * B + ln] = (int16_t)iter;
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
          if ((i >= 1 || a.mode == MODE_CONTINUOUS) && success_cond(c, L.xtrue)) L.succ = 1;
          for (int k = 0; k < 4; ++k) L.xfin[k] = L.xtrue[k];
          if (a.mode == MODE_CONTINUOUS) {
            // trajectorySimulateC.py:325-409, as post_kernel (sim.cuh): the sample substep sees the previous command
            const double nm = c.mean_mtn, hh = a.T_cont;
            for (int k = 0; k < 4; ++k) xn[k] = L.xtrue[k];
            const int nr = min(L.sub / a.noise_hold_sub, a.n_refresh - 1);
            const double w0 = a.noise_in ? a.noise_in[((size_t)nr * 2 + 0) * B + ln] : 0.0;
            const double w1 = a.noise_in ? a.noise_in[((size_t)nr * 2 + 1) * B + ln] : 0.0;
            if (!c.delta_v) {
              rk4_substep(xn, uprev[0], uprev[1], nm, hh);
            } else {
              rk4_substep(xn, 0.0, 0.0, nm, hh);
              xn[2] += uprev[0];
              xn[3] += uprev[1];
            }
            xn[0] += w0;
            xn[1] += w1;
          } else {
            plant_lin(c, L.xtrue, uprev, L.noise, xn);
          }
          double xe[6];
          if (c.has_noise) {
            if (!estimator_step(c, L.ux, L.uP, uprev, xn)) L.ukf_clamp = 1;
            for (int k = 0; k < 6; ++k) xe[k] = L.ux[k];
          } else {
            for (int k = 0; k < 4; ++k) xe[k] = xn[k];
            xe[4] = xe[5] = 0.0;
          }
          gen_geometry(a, L, xe);
          if (c.in_track) {
            const double t = xe[0];
            xe[0] = xe[1];
            xe[1] = t;
          }
          for (int k = 0; k < 6; ++k) L.xstore[k] = xe[k];
          if (a.out.x_est) for (int k = 0; k < 6; ++k) a.out.x_est[((size_t)k * T1 + i + 1) * B + ln] = xe[k];
          if (a.out.x_true) for (int k = 0; k < 4; ++k) a.out.x_true[((size_t)k * T1 + i + 1) * B + ln] = xn[k];
          if (a.mode == MODE_CONTINUOUS) {
            const size_t NS = a.out.NS;
            const double nm = c.mean_mtn, hh = a.T_cont;
            int sub = L.sub;
            auto put_sub = [&](int sdx) {
              if (a.out.x_true_sub) for (int k = 0; k < 4; ++k) a.out.x_true_sub[((size_t)k * NS + sdx + 1) * B + ln] = xn[k];
              if (a.out.ctrl_sub) {
                a.out.ctrl_sub[((size_t)0 * NS + sdx + 1) * B + ln] = u[0];
                a.out.ctrl_sub[((size_t)1 * NS + sdx + 1) * B + ln] = u[1];
              }
              if (a.out.ctrlr_sub) a.out.ctrlr_sub[(size_t)sdx * B + ln] = (uint8_t)code;
            };
            put_sub(sub);
            sub += 1;
            const int next_sample = (i + 2) * a.ratio;
            const bool more_samples = (i + 2) < a.nsteps;
            while (true) {
              if (sub >= a.n_sub_total - 1) {
                if (success_cond(c, xn)) L.succ = 1;
                for (int k = 0; k < 4; ++k) L.xfin[k] = xn[k];
                L.fin = 1;
                break;
              }
              if (terminated(c, xn)) { L.iterm = sub; L.fin = 1; break; }
              if (more_samples && sub == next_sample) break;
              if (success_cond(c, xn)) L.succ = 1;
              for (int k = 0; k < 4; ++k) L.xfin[k] = xn[k];
              const int nr = min(sub / a.noise_hold_sub, a.n_refresh - 1);
              const double w0 = a.noise_in ? a.noise_in[((size_t)nr * 2 + 0) * B + ln] : 0.0;
              const double w1 = a.noise_in ? a.noise_in[((size_t)nr * 2 + 1) * B + ln] : 0.0;
              if (!c.delta_v) rk4_substep(xn, u[0], u[1], nm, hh);
              else rk4_substep(xn, 0.0, 0.0, nm, hh);
              xn[0] += w0;
              xn[1] += w1;
              put_sub(sub);
              sub += 1;
            }
            L.sub = sub;
            for (int k = 0; k < 4; ++k) L.xtrue[k] = xn[k];
            L.step = i + 1;
          } else {
          for (int k = 0; k < 4; ++k) L.xtrue[k] = xn[k];
          if (c.has_noise && ((i + 1) % c.noise_length == 0)) {
            const int r = min((i + 1) / c.noise_length, a.n_refresh - 1);
            L.noise[0] = a.noise_in[((size_t)r * 2 + 0) * B + ln];
            L.noise[1] = a.noise_in[((size_t)r * 2 + 1) * B + ln];
          }
          L.step = i + 1;
          if (i + 1 >= a.nsteps) L.fin = 1;
          else if (terminated(c, xn)) { L.iterm = i + 1; L.fin = 1; }
          }
        }
      }
      __syncthreads();
      GP_MARK(5)
    }

    // ---------------- lane done
    for (int j = tid; j < n; j += T) a.xs[(size_t)ln * n + j] = x[j];
    for (int i = tid; i < m; i += T) {
      a.zs[(size_t)ln * m + i] = z[i];
      a.ys[(size_t)ln * m + i] = y[i];
    }
    if (tid == 0) {
      a.rho[ln] = L.rho;
      if (a.mode == MODE_QP_ONLY) {
        a.status[ln] = L.status;
        a.iter[ln] = L.iter;
        a.u0[ln] = L.u0[0];
        a.u0[B + ln] = L.u0[1];
        atomicAdd(&a.tot[1], 1ull);
      } else {
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
        const double f = (fd == fd) ? fd : 0.0;
        if (a.out.fd_all) a.out.fd_all[ln] = f;   // stats[0], stats[1]: fixed-order sum afterwards (stats_fd_kernel)
        if (L.succ) atomicAdd(&a.stats[2], 1.0);
        atomicAdd(&a.stats[3], 1.0);
        atomicAdd(&a.stats[4], (double)L.iterm);
        atomicAdd(&a.stats[5], (double)L.nsolve);
        if (L.ukf_clamp) atomicAdd(&a.stats[8], 1.0);
        if (L.iterm < ((a.mode == MODE_CONTINUOUS) ? a.n_sub_total : a.nsteps)) atomicAdd(&a.stats[9], 1.0);
        atomicAdd(&a.tot[1], (unsigned long long)L.nsolve);
      }
    }
  }
  GP_FLUSH
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(512));
  if (tid == 0) {
    if (my_iters) atomicAdd(&a.tot[0], my_iters);
    if (my_rebuilds) atomicAdd(&a.tot[2], my_rebuilds);
  }
}
