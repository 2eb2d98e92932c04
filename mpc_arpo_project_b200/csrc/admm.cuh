// ADMM block kernel: one check_termination period (25 iterations) of OSQP's ADMM for every
// live lane, then OSQP's termination / infeasibility / adaptive-rho logic.
//
// Replaces  res = prob.solve()  (reference src/trajectorySimulate.py:296; OSQP 0.6.x osqp_solve,
// restated in oracle/osqp_ref.py) in the reduced-KKT + spectral form of oracle/batched_ref.py:
//     r  = sigma*x - q + A'(rho_vec.*z - y)
//     xt = V diag(1/(1+rho*lam)) V' r            (two dense n x n mat-vecs, V shared by the CTA)
//     zt = A xt ; x+ = a*xt+(1-a)*x ; zr = a*zt+(1-a)*z
//     z+ = clip(zr + y./rho_vec, l, u) ; y+ = y + rho_vec.*(zr - z+)
// Mapping: one warp per trajectory ("lane"); its iterates x (n), z, y (m) live in registers,
// element i on thread i%32.  A CTA holds lanes of ONE sign variant, so the variant's constant
// blob (V, grouped-ELL A / A' / P, scalings, bound templates) is staged once into shared memory
// with a bulk async copy (TMA engine) and shared by all its warps.
#pragma once
#include "common.cuh"

template <int NS>
__device__ __forceinline__ void ell_apply(const double *__restrict__ vals, const uint16_t *__restrict__ cols,
                                          const int2 *__restrict__ grp, int lane, const double *__restrict__ vec,
                                          double (&out)[NS]) {
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int2 g = grp[s];
    const double *v = vals + g.x + lane;
    const uint16_t *c = cols + g.x + lane;
    double acc = 0.0;
    for (int e = 0; e < g.y; ++e) acc = fma(v[e * 32], vec[c[e * 32]], acc);
    out[s] = acc;
  }
}

template <int NXS, int MZS, bool VSMEM, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) admm_block_kernel(const __grid_constant__ AdmmArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int W = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = a.n, m = a.m;

  // ---- which (variant, tile) is this CTA
  int b = blockIdx.x, v = 0, cnt_v = 0;
  for (; v < 4; ++v) {
    cnt_v = a.cnt[v];
    const int nt = (cnt_v + W - 1) / W;
    if (b < nt) break;
    b -= nt;
  }
  if (v == 4) return;

  // ---- stage the variant's constant blob into shared memory (one elected thread drives the TMA engine)
  const BlobHdr &h = a.hdr;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + h.total);
  double *scratch = reinterpret_cast<double *>(smem + h.total + 16);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(bar, (uint32_t)h.total);
    const unsigned char *src = a.blob[v];
    for (int off = 0; off < h.total; off += 32768) {
      const int nb = min(32768, h.total - off);
      bulk_g2s(smem + off, src + off, (uint32_t)nb, bar);
    }
  }
  __syncthreads();
  mbar_wait(bar, 0);

  const double *lam = reinterpret_cast<const double *>(smem + h.off_lam);
  const double *qv = reinterpret_cast<const double *>(smem + h.off_q);
  const double *Dv = reinterpret_cast<const double *>(smem + h.off_D);
  const double *Dinv = reinterpret_cast<const double *>(smem + h.off_Dinv);
  const double *Ev = reinterpret_cast<const double *>(smem + h.off_E);
  const double *Einv = reinterpret_cast<const double *>(smem + h.off_Einv);
  const double *lt = reinterpret_cast<const double *>(smem + h.off_lt);
  const double *ut = reinterpret_cast<const double *>(smem + h.off_ut);
  const double *Avals = reinterpret_cast<const double *>(smem + h.off_Av);
  const double *ATvals = reinterpret_cast<const double *>(smem + h.off_ATv);
  const double *Pvals = reinterpret_cast<const double *>(smem + h.off_Pv);
  const uint16_t *Acols = reinterpret_cast<const uint16_t *>(smem + h.off_Ac);
  const uint16_t *ATcols = reinterpret_cast<const uint16_t *>(smem + h.off_ATc);
  const uint16_t *Pcols = reinterpret_cast<const uint16_t *>(smem + h.off_Pc);
  const int2 *Agrp = reinterpret_cast<const int2 *>(smem + h.off_Ag);
  const int2 *ATgrp = reinterpret_cast<const int2 *>(smem + h.off_ATg);
  const int2 *Pgrp = reinterpret_cast<const int2 *>(smem + h.off_Pg);
  const uint8_t *flags = reinterpret_cast<const uint8_t *>(smem + h.off_flags);
  const double *V = VSMEM ? reinterpret_cast<const double *>(smem + h.off_V) : a.Vg[v];

  double *bm = scratch + (size_t)warp * (m + 2 * n);
  double *bn = bm + m;
  double *bn2 = bn + n;

  const int pos = b * W + warp;
  unsigned long long my_iters = 0;
  if (pos < cnt_v) {
    const int ln = a.list[(size_t)v * a.B + pos];
    const int B = a.B;
    double rho = a.rho[ln];
    int iter = a.iter[ln];
    const double xh0 = a.par[0 * (size_t)B + ln], xh1 = a.par[1 * (size_t)B + ln], xh2 = a.par[2 * (size_t)B + ln],
                 xh3 = a.par[3 * (size_t)B + ln], val = a.par[4 * (size_t)B + ln], dp0 = a.par[5 * (size_t)B + ln],
                 dp1 = a.par[6 * (size_t)B + ln];

    double x[NXS], dsc[NXS], z[MZS], y[MZS], lo[MZS], hi[MZS], rv[MZS], rinv[MZS], dy[MZS];
    const double *xg = a.xs + (size_t)ln * n, *zg = a.zs + (size_t)ln * m, *yg = a.ys + (size_t)ln * m;
    int flip = 0;
#pragma unroll
    for (int s = 0; s < NXS; ++s) {
      const int j = lane + 32 * s;
      x[s] = (j < n) ? xg[j] : 0.0;
      dsc[s] = (j < n) ? 1.0 / (1.0 + rho * lam[j]) : 0.0;
    }
#pragma unroll
    for (int s = 0; s < MZS; ++s) {
      const int i = lane + 32 * s;
      const bool ok = i < m;
      z[s] = ok ? zg[i] : 0.0;
      y[s] = ok ? yg[i] : 0.0;
      double l_ = ok ? lt[i] : 0.0, u_ = ok ? ut[i] : 0.0;
      const uint8_t f = ok ? flags[i] : (uint8_t)8;
      if (ok) {
        if (i < 4) {
          const double xh = (i == 0) ? xh0 : (i == 1) ? xh1 : (i == 2) ? xh2 : xh3;
          l_ = u_ = -xh * Ev[i];
        } else if (i >= m - 2) {
          l_ = u_ = ((i == m - 2) ? dp0 : dp1) * Ev[i];
        } else if (i >= a.nX && i < a.nX + 5 * (a.Nb + 1) && (i - a.nX) % 5 == 3) {
          u_ = val * Ev[i];
          if (u_ - l_ < MPCB_RHO_TOL) flip = 1;
        }
      }
      lo[s] = l_;
      hi[s] = u_;
      rv[s] = (f & 8) ? MPCB_RHO_MIN : ((f & 4) ? MPCB_RHO_EQ * rho : rho);
      rinv[s] = 1.0 / rv[s];
      dy[s] = 0.0;
    }
    if (__any_sync(0xffffffffu, flip) && lane == 0) a.flip[ln] = 1;

    const double sigma = a.sigma, alpha = a.alpha, oma = 1.0 - a.alpha;
    const double *vrow[NXS];
#pragma unroll
    for (int s = 0; s < NXS; ++s) vrow[s] = V + (size_t)min(lane + 32 * s, n - 1) * n;

    // =========================== check_every ADMM iterations ===========================
    for (int it = 0; it < a.check_every; ++it) {
#pragma unroll
      for (int s = 0; s < MZS; ++s) {
        const int i = lane + 32 * s;
        if (i < m) bm[i] = rv[s] * z[s] - y[s];
      }
      __syncwarp();
      {
        double acc[NXS];
        ell_apply<NXS>(ATvals, ATcols, ATgrp, lane, bm, acc);
#pragma unroll
        for (int s = 0; s < NXS; ++s) {
          const int j = lane + 32 * s;
          if (j < n) bn[j] = sigma * x[s] - qv[j] + acc[s];
        }
      }
      __syncwarp();
      {  // w = dsc .* (V' r)
        double acc[NXS];
#pragma unroll
        for (int s = 0; s < NXS; ++s) acc[s] = 0.0;
        const double *vk = V + lane;
#pragma unroll 4
        for (int k = 0; k < n; ++k) {
          const double rk = bn[k];
#pragma unroll
          for (int s = 0; s < NXS; ++s) acc[s] = fma(VSMEM ? vk[32 * s] : __ldg(vk + 32 * s), rk, acc[s]);
          vk += n;
        }
#pragma unroll
        for (int s = 0; s < NXS; ++s) {
          const int j = lane + 32 * s;
          if (j < n) bn2[j] = acc[s] * dsc[s];
        }
      }
      __syncwarp();
      {  // xt = V w
        double acc[NXS];
#pragma unroll
        for (int s = 0; s < NXS; ++s) acc[s] = 0.0;
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
          const double wj = bn2[j];
#pragma unroll
          for (int s = 0; s < NXS; ++s) acc[s] = fma(VSMEM ? vrow[s][j] : __ldg(vrow[s] + j), wj, acc[s]);
        }
#pragma unroll
        for (int s = 0; s < NXS; ++s) {
          const int i = lane + 32 * s;
          if (i < n) bn[i] = acc[s];
        }
      }
      __syncwarp();
      {
        double zt[MZS];
        ell_apply<MZS>(Avals, Acols, Agrp, lane, bn, zt);
#pragma unroll
        for (int s = 0; s < MZS; ++s) {
          const double zr = alpha * zt[s] + oma * z[s];
          const double zn = fmin(fmax(zr + rinv[s] * y[s], lo[s]), hi[s]);
          dy[s] = rv[s] * (zr - zn);
          y[s] += dy[s];
          z[s] = zn;
        }
#pragma unroll
        for (int s = 0; s < NXS; ++s) {
          const int j = lane + 32 * s;
          if (j < n) x[s] = alpha * bn[j] + oma * x[s];
        }
      }
      __syncwarp();
    }
    iter += a.check_every;
    my_iters = (unsigned long long)a.check_every;

    // =========================== update_info (OSQP auxil.c) ===========================
#pragma unroll
    for (int s = 0; s < NXS; ++s) {
      const int j = lane + 32 * s;
      if (j < n) bn[j] = x[s];
    }
#pragma unroll
    for (int s = 0; s < MZS; ++s) {
      const int i = lane + 32 * s;
      if (i < m) bm[i] = y[s];
    }
    __syncwarp();
    double Ax[MZS], Px[NXS], Aty[NXS];
    ell_apply<MZS>(Avals, Acols, Agrp, lane, bn, Ax);
    ell_apply<NXS>(Pvals, Pcols, Pgrp, lane, bn, Px);
    ell_apply<NXS>(ATvals, ATcols, ATgrp, lane, bm, Aty);
    double pri_u = 0, nz_u = 0, nax_u = 0, pri_s = 0, nz_s = 0, nax_s = 0;
#pragma unroll
    for (int s = 0; s < MZS; ++s) {
      const int i = lane + 32 * s;
      if (i < m) {
        const double ei = Einv[i], pv = Ax[s] - z[s];
        pri_u = fmax(pri_u, fabs(ei * pv));
        nz_u = fmax(nz_u, fabs(ei * z[s]));
        nax_u = fmax(nax_u, fabs(ei * Ax[s]));
        pri_s = fmax(pri_s, fabs(pv));
        nz_s = fmax(nz_s, fabs(z[s]));
        nax_s = fmax(nax_s, fabs(Ax[s]));
      }
    }
    double dua_u = 0, npx_u = 0, naty_u = 0, dua_s = 0, npx_s = 0, naty_s = 0;
#pragma unroll
    for (int s = 0; s < NXS; ++s) {
      const int j = lane + 32 * s;
      if (j < n) {
        const double di = Dinv[j], dv = qv[j] + Px[s] + Aty[s];
        dua_u = fmax(dua_u, fabs(di * dv));
        npx_u = fmax(npx_u, fabs(di * Px[s]));
        naty_u = fmax(naty_u, fabs(di * Aty[s]));
        dua_s = fmax(dua_s, fabs(dv));
        npx_s = fmax(npx_s, fabs(Px[s]));
        naty_s = fmax(naty_s, fabs(Aty[s]));
      }
    }
    pri_u = warp_max(pri_u); nz_u = warp_max(nz_u); nax_u = warp_max(nax_u);
    pri_s = warp_max(pri_s); nz_s = warp_max(nz_s); nax_s = warp_max(nax_s);
    dua_u = warp_max(dua_u) * a.cinv; npx_u = warp_max(npx_u); naty_u = warp_max(naty_u);
    dua_s = warp_max(dua_s); npx_s = warp_max(npx_s); naty_s = warp_max(naty_s);

    // =========================== check_termination ===========================
    // primal-infeasibility certificate pieces (is_primal_infeasible): computed lazily
    bool cert_ready = false;
    double ndy = 0, lhs = 0, natdy = 0;
    auto check = [&](double k) -> int {
      const double eps_p = k * a.eps_abs + k * a.eps_rel * fmax(nz_u, nax_u);
      const double eps_d = k * a.eps_abs + k * a.eps_rel * a.cinv * fmax(a.qn_unscaled, fmax(naty_u, npx_u));
      const bool prim_ok = pri_u < eps_p, dual_ok = dua_u < eps_d;
      if (prim_ok && dual_ok) return (k > 1.0) ? 2 : 1;
      if (!prim_ok) {
        if (!cert_ready) {
          double n1 = 0, l1 = 0;
#pragma unroll
          for (int s = 0; s < MZS; ++s) {
            const int i = lane + 32 * s;
            if (i < m) {
              const uint8_t f = flags[i];
              double d = dy[s];
              if ((f & 3) == 3) d = 0.0;
              else if (f & 2) d = fmin(d, 0.0);
              else if (f & 1) d = fmax(d, 0.0);
              dy[s] = d;
              bm[i] = d;
              n1 = fmax(n1, fabs(Ev[i] * d));
              l1 += hi[s] * fmax(d, 0.0) + lo[s] * fmin(d, 0.0);
            }
          }
          ndy = warp_max(n1);
          lhs = warp_sum(l1);
          __syncwarp();
          double atdy[NXS];
          ell_apply<NXS>(ATvals, ATcols, ATgrp, lane, bm, atdy);
          double n2 = 0;
#pragma unroll
          for (int s = 0; s < NXS; ++s) {
            const int j = lane + 32 * s;
            if (j < n) n2 = fmax(n2, fabs(Dinv[j] * atdy[s]));
          }
          natdy = warp_max(n2);
          cert_ready = true;
        }
        const double eps_i = k * a.eps_pinf;
        if (ndy > MPCB_DIV_TOL && lhs < -eps_i * ndy && natdy < eps_i * ndy) return (k > 1.0) ? 3 : -3;
      }
      return -10;
    };
    int st = check(1.0);
    if (st == -10) {
      if (a.adaptive && (iter % a.adapt_interval == 0)) {
        // compute_rho_estimate on the SCALED residual vectors (OSQP 0.6.x)
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
    double *xo = a.xs + (size_t)ln * n, *zo = a.zs + (size_t)ln * m, *yo = a.ys + (size_t)ln * m;
#pragma unroll
    for (int s = 0; s < NXS; ++s) {
      const int j = lane + 32 * s;
      if (j < n) xo[j] = x[s];
    }
#pragma unroll
    for (int s = 0; s < MZS; ++s) {
      const int i = lane + 32 * s;
      if (i < m) {
        zo[i] = z[s];
        yo[i] = y[s];
      }
    }
    if (lane == 0) {
      a.rho[ln] = rho;
      a.iter[ln] = iter;
      a.status[ln] = st;
      if (st != -10) {
        a.u0[ln] = Dv[a.uoff] * bn[a.uoff];
        a.u0[(size_t)B + ln] = Dv[a.uoff + 1] * bn[a.uoff + 1];
        a.lane_state[ln] = LANE_SOLVE_DONE;
      }
    }
  }
  // one global atomic per CTA for the iteration counter
  __shared__ unsigned long long cta_iters;
  if (threadIdx.x == 0) cta_iters = 0;
  __syncthreads();
  if (lane == 0 && my_iters) atomicAdd(&cta_iters, my_iters);
  __syncthreads();
  if (threadIdx.x == 0 && cta_iters) atomicAdd(a.iter_total, cta_iters);
}
