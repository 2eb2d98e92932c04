// Per-lane closed-loop kernels: controller selection + norm clip, plant step (linear CW or RK4
// nonlinear), unscented Kalman / disturbance estimator, QP parameter refresh, lane binning.
// One thread per trajectory; all per-lane state is SoA [field][B] so accesses coalesce.
//
// Reference lines each piece stands in for (src/trajectorySimulate.py unless noted):
//   select_control  :299-319      plant_lin  :323-324       ukf_step  :121-130,:329-337
//   qp_params       :340-348 + src/simhelpers.py:66-67,124,137-138
//   rk4_substep     src/trajectorySimulateC.py:64-79,372-380 (solve_ivp -> fixed-step RK4)
#pragma once
#include "common.cuh"

__device__ __forceinline__ void plant_lin(const SimConst &c, const double *x, const double *u, const double *w, double *xn) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc += c.Ad[i * 4 + j] * x[j];
    acc += c.Bd[i * 2 + 0] * u[0] + c.Bd[i * 2 + 1] * u[1];
    xn[i] = acc + (i < 2 ? w[i] : 0.0);
  }
}

__device__ __forceinline__ void state_eqn_n(const double *x, double u0, double u1, double nm, double *dx) {
  const double R_T = 500e+03 + 6378.1e+03;
  const double mu = (nm * nm) * (R_T * R_T * R_T);
  const double rx = R_T + x[0];
  const double r2 = rx * rx + x[1] * x[1];
  // mu / r^3 through ONE reciprocal square root instead of a square root and two divisions: this function is evaluated
  // 2000 times per control step and lane (RK4 at 1 ms), on a single dependent chain; 1-2 ulp from (mu * rx) / r3
  const double rs = rsqrt(r2);
  const double mur3 = mu * (rs * rs * rs);
  dx[0] = x[2];
  dx[1] = x[3];
  dx[2] = 2 * nm * x[3] + (nm * nm) * x[0] - mur3 * rx + mu / (R_T * R_T) + u0;
  dx[3] = -2 * nm * x[2] + (nm * nm) * x[1] - mur3 * x[1] + u1;
}

__device__ __forceinline__ void rk4_substep(double *x, double u0, double u1, double nm, double h) {
  double k1[4], k2[4], k3[4], k4[4], t[4];
  state_eqn_n(x, u0, u1, nm, k1);
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = x[i] + 0.5 * h * k1[i];
  state_eqn_n(t, u0, u1, nm, k2);
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = x[i] + 0.5 * h * k2[i];
  state_eqn_n(t, u0, u1, nm, k3);
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = x[i] + h * k3[i];
  state_eqn_n(t, u0, u1, nm, k4);
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = x[i] + (h / 6.0) * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
}

// Merwe scaled sigma points n=6, alpha=.1, beta=2, kappa=-1 (ref :130): lambda+n = 0.05
#define UKF_NPL 0.05
#define UKF_WM0 (-5.95 / 0.05)
#define UKF_WC0 (-5.95 / 0.05 + (1.0 - 0.01 + 2.0))
#define UKF_WI (0.5 / 0.05)

// upper Cholesky U'U = s*P (scipy.linalg.cholesky default, filterpy's `sqrt`).  With R = 0 the
// posterior covariance is singular in the measured directions and rounding can leave a pivot at
// -1e-18: the reference then dies with LinAlgError.  The engine continues with the positive
// SEMI-definite factor instead (pivot <= 0 -> that row of U is zero) and reports the lane
// (returns false); a pivot of +1e-18 gives the same factor to ~1e-9.
__device__ __forceinline__ bool chol_upper6(const double *P, double s, double *U) {
  bool ok = true;
  for (int i = 0; i < 6; ++i) {
    for (int j = 0; j < 6; ++j) U[i * 6 + j] = 0.0;
  }
  for (int i = 0; i < 6; ++i) {
    double d = s * P[i * 6 + i];
    for (int k = 0; k < i; ++k) d -= U[k * 6 + i] * U[k * 6 + i];
    if (!(d > 0.0)) {
      ok = false;
      continue;                               // row i stays zero
    }
    const double r = sqrt(d);
    U[i * 6 + i] = r;
    for (int j = i + 1; j < 6; ++j) {
      double v = s * P[i * 6 + j];
      for (int k = 0; k < i; ++k) v -= U[k * 6 + i] * U[k * 6 + j];
      U[i * 6 + j] = v / r;
    }
  }
  return ok;
}

__device__ __forceinline__ bool sigma_points(const double *x, const double *P, double *sig /*[13][6]*/) {
  double U[36];
  const bool ok = chol_upper6(P, UKF_NPL, U);
  for (int j = 0; j < 6; ++j) sig[j] = x[j];
  for (int k = 0; k < 6; ++k)
    for (int j = 0; j < 6; ++j) {
      sig[(k + 1) * 6 + j] = x[j] + U[k * 6 + j];
      sig[(k + 7) * 6 + j] = x[j] - U[k * 6 + j];
    }
  return ok;
}

// kf.predict(u); kf.update(z) of filterpy 1.4.5 as restated in oracle/ukf_ref.py (R = 0).
// Returns false when a Cholesky pivot had to be clamped (the reference would have raised).
__device__ __noinline__ bool ukf_step(const SimConst &c, double *x, double *P, const double *u, const double *zmeas) {
  double sig[78], sf[78];
  bool ok = sigma_points(x, P, sig);
  for (int k = 0; k < 13; ++k)
    for (int i = 0; i < 6; ++i) {
      double acc = 0.0;
      for (int j = 0; j < 6; ++j) acc += c.Ao[i * 6 + j] * sig[k * 6 + j];
      sf[k * 6 + i] = acc + c.Bou[i * 2] * u[0] + c.Bou[i * 2 + 1] * u[1];
    }
  double xm[6], Pm[36];
  for (int i = 0; i < 6; ++i) {
    double acc = UKF_WM0 * sf[i];
    for (int k = 1; k < 13; ++k) acc += UKF_WI * sf[k * 6 + i];
    xm[i] = acc;
  }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double acc = 0.0;
      for (int k = 0; k < 13; ++k) acc += (k == 0 ? UKF_WC0 : UKF_WI) * (sf[k * 6 + i] - xm[i]) * (sf[k * 6 + j] - xm[j]);
      Pm[i * 6 + j] = acc + c.Qw[i * 6 + j];
    }
  ok &= sigma_points(xm, Pm, sf);           // filterpy 1.4.5 regenerates the points after predict
  double zs[26], zp[2] = {0.0, 0.0};
  for (int k = 0; k < 13; ++k) {
    const double a = sf[k * 6], b = sf[k * 6 + 1];
    zs[k * 2] = sqrt(a * a + b * b);
    zs[k * 2 + 1] = atan2(b, a);
    const double w = (k == 0 ? UKF_WM0 : UKF_WI);
    zp[0] += w * zs[k * 2];
    zp[1] += w * zs[k * 2 + 1];
  }
  double S[4] = {0, 0, 0, 0}, Pxz[12];
  for (int i = 0; i < 12; ++i) Pxz[i] = 0.0;
  for (int k = 0; k < 13; ++k) {
    const double w = (k == 0 ? UKF_WC0 : UKF_WI);
    const double d0 = zs[k * 2] - zp[0], d1 = zs[k * 2 + 1] - zp[1];
    S[0] += w * d0 * d0; S[1] += w * d0 * d1; S[2] += w * d1 * d0; S[3] += w * d1 * d1;
    for (int i = 0; i < 6; ++i) {
      const double dx = sf[k * 6 + i] - xm[i];
      Pxz[i * 2] += w * dx * d0;
      Pxz[i * 2 + 1] += w * dx * d1;
    }
  }
  const double det = S[0] * S[3] - S[1] * S[2];
  const double SI[4] = {S[3] / det, -S[1] / det, -S[2] / det, S[0] / det};
  double K[12];
  for (int i = 0; i < 6; ++i) {
    K[i * 2] = Pxz[i * 2] * SI[0] + Pxz[i * 2 + 1] * SI[2];
    K[i * 2 + 1] = Pxz[i * 2] * SI[1] + Pxz[i * 2 + 1] * SI[3];
  }
  const double y0 = zmeas[0] - zp[0], y1 = zmeas[1] - zp[1];
  for (int i = 0; i < 6; ++i) x[i] = xm[i] + K[i * 2] * y0 + K[i * 2 + 1] * y1;
  for (int i = 0; i < 6; ++i) {
    const double ks0 = K[i * 2] * S[0] + K[i * 2 + 1] * S[2], ks1 = K[i * 2] * S[1] + K[i * 2 + 1] * S[3];
    for (int j = 0; j < 6; ++j) P[i * 6 + j] = Pm[i * 6 + j] - (ks0 * K[j * 2] + ks1 * K[j * 2 + 1]);
  }
  return ok;
}

__device__ __forceinline__ bool terminated(const SimConst &c, const double *x) {
  const double r = sqrt(x[0] * x[0] + x[1] * x[1]);
  return (r < c.r_p) || ((c.in_track ? x[1] : x[0]) < c.r_p - c.r_tol);
}

__device__ __forceinline__ bool success_cond(const SimConst &c, const double *x) {
  const double dx = x[0] - c.xr[0], dy = x[1] - c.xr[1];
  const double dist = sqrt(dx * dx + dy * dy);
  if (!(dist <= c.suc_dist)) return false;
  const double ang = fabs(atan(x[3] / x[2])) * (180.0 / 3.141592653589793);
  return ang <= c.suc_ang_deg;
}

// QP parameters the two prob.update calls depend on (trajectorySimulate.py:340-348).
__device__ __forceinline__ int write_qp_params(const PostArgs &a, int ln, const double *xe /*6, unswapped*/) {
  const size_t B = a.B;
  a.par[0 * B + ln] = xe[0];
  a.par[1 * B + ln] = xe[1];
  a.par[2 * B + ln] = xe[2];
  a.par[3 * B + ln] = xe[3];
  a.par[4 * B + ln] = fabs(xe[0] - a.sc.xr[0]) + fabs(xe[1] - a.sc.xr[1]);
  a.par[5 * B + ln] = a.sc.is_reject ? xe[4] : 0.0;
  a.par[6 * B + ln] = a.sc.is_reject ? xe[5] : 0.0;
  return (xe[2] >= 0 ? 0 : 1) + (xe[3] >= 0 ? 0 : 2);      // sign(0) = +1, simhelpers.py:66-67
}

// Append live lanes to the next round's per-variant lists (warp-aggregated atomics).
// `fresh`: the lane starts a NEW solve (its parameters were just refreshed) -- the moment a lane may be deferred.
__device__ __forceinline__ void bin_lane(const PostArgs &a, int ln, bool live, int variant, bool fresh = false) {
  const unsigned lane = threadIdx.x & 31u;
  const bool defer = live && fresh && a.defer_below > 0.0 && a.par[(size_t)4 * a.B + ln] < a.defer_below;
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const unsigned mask = __ballot_sync(0xffffffffu, live && !defer && variant == v);
    if (mask) {
      int base = 0;
      const int leader = __ffs(mask) - 1;
      if ((int)lane == leader) base = atomicAdd(&a.cnt_next[v], __popc(mask));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (live && !defer && variant == v) a.list_next[(size_t)v * a.B + base + __popc(mask & ((1u << lane) - 1u))] = ln;
    }
  }
  if (__any_sync(0xffffffffu, defer) && defer) {        // rare: plain atomics
    a.list_def[(size_t)variant * a.B + atomicAdd(&a.cnt_def[variant], 1)] = ln;
    a.lane_state[ln] = LANE_DEFERRED;                   // post_kernel re-bins LANE_SOLVING lanes every round: not this one
  }
}

// Controller selection + sequential norm clip (trajectorySimulate.py:299-319; no-debris path).
__device__ __forceinline__ int select_control(const PostArgs &a, int ln, int status, const double *xs, double *u, double *uraw) {
  const size_t B = a.B;
  int code;
  if (status != 1) {
    const double xi = a.ls.xintf[ln] + xs[0] - a.sc.xr[0];
    a.ls.xintf[ln] = xi;
    for (int r = 0; r < 2; ++r) {
      double acc = 0.0;
      for (int j = 0; j < 4; ++j) acc += a.sc.Kpf[r * 4 + j] * xs[j];
      u[r] = -acc - a.sc.Kif[r] * xi;
    }
    code = 2;
  } else {
    a.ls.xintf[ln] = 0.0;
    u[0] = a.u0[ln];
    u[1] = a.u0[B + ln];
    code = 1;
  }
  uraw[0] = u[0];
  uraw[1] = u[1];
  const double nrm = sqrt(u[0] * u[0] + u[1] * u[1]);
  if (nrm > a.sc.umax0) {
    u[0] = u[0] * (a.sc.umax0 / nrm);
    const double nrm2 = sqrt(u[0] * u[0] + u[1] * u[1]);
    u[1] = u[1] * (a.sc.umax0 / nrm2);
  }
  return code;
}

// Linear Kalman filter on POSITION measurements, R = 0 -- the estimator of the reference's prototype
// misc/MPCrendezKALMANdisturb.py:261-266 (restated in oracle/kf_ref.py), called where the simulators call the UKF:
//   x- = Ao x + Bou u;  P- = Ao P Ao' + Qw;  L = P- Co' (Co P- Co')^-1;  x = x- + L (y - Co x-);  P = (I - L Co) P-
// with Co = [I2 0].  Always returns true (no Cholesky to clamp).
__device__ __noinline__ bool kf_step(const SimConst &c, double *x, double *P, const double *u, const double *ymeas) {
  double xm[6], AP[36], Pm[36];
  for (int i = 0; i < 6; ++i) {
    double acc = 0.0;
    for (int j = 0; j < 6; ++j) acc += c.Ao[i * 6 + j] * x[j];
    xm[i] = acc + c.Bou[i * 2] * u[0] + c.Bou[i * 2 + 1] * u[1];
  }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double acc = 0.0;
      for (int k = 0; k < 6; ++k) acc += c.Ao[i * 6 + k] * P[k * 6 + j];
      AP[i * 6 + j] = acc;
    }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double acc = 0.0;
      for (int k = 0; k < 6; ++k) acc += AP[i * 6 + k] * c.Ao[j * 6 + k];
      Pm[i * 6 + j] = acc + c.Qw[i * 6 + j];
    }
  const double s00 = Pm[0], s01 = Pm[1], s10 = Pm[6], s11 = Pm[7];
  const double det = s00 * s11 - s01 * s10;
  const double i00 = s11 / det, i01 = -s01 / det, i10 = -s10 / det, i11 = s00 / det;
  double Lg[12];
  for (int i = 0; i < 6; ++i) {
    Lg[i * 2] = Pm[i * 6] * i00 + Pm[i * 6 + 1] * i10;
    Lg[i * 2 + 1] = Pm[i * 6] * i01 + Pm[i * 6 + 1] * i11;
  }
  const double r0 = ymeas[0] - xm[0], r1 = ymeas[1] - xm[1];
  for (int i = 0; i < 6; ++i) x[i] = xm[i] + Lg[i * 2] * r0 + Lg[i * 2 + 1] * r1;
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) P[i * 6 + j] = Pm[i * 6 + j] - (Lg[i * 2] * Pm[j] + Lg[i * 2 + 1] * Pm[6 + j]);
  return true;
}

// The estimator the problem selects; zsrc = true state the measurement is taken from.
__device__ __forceinline__ bool estimator_step(const SimConst &c, double *x, double *P, const double *u, const double *xn) {
  if (c.estimator == MPCB_EST_KF) {
    const double y[2] = {xn[0], xn[1]};
    return kf_step(c, x, P, u, y);
  }
  const double zm[2] = {sqrt(xn[0] * xn[0] + xn[1] * xn[1]), atan2(xn[1], xn[0])};
  return ukf_step(c, x, P, u, zm);
}

// Estimator + QP refresh shared by both simulators (trajectorySimulate.py:329-348).
// xn: true state the measurement is taken from; uprev: ctrls[:, i]; est_idx: telemetry column.
__device__ __forceinline__ int estimate_and_refresh(const PostArgs &a, int ln, const double *xn, const double *uprev, int est_idx) {
  const size_t B = a.B;
  double xe[6];
  if (a.sc.has_noise) {
    double ux[6], uP[36];
    for (int i = 0; i < 6; ++i) ux[i] = a.ls.ux[i * B + ln];
    for (int i = 0; i < 36; ++i) uP[i] = a.ls.uP[i * B + ln];
    if (!estimator_step(a.sc, ux, uP, uprev, xn)) a.ls.ukf_clamp[ln] = 1;
    for (int i = 0; i < 6; ++i) a.ls.ux[i * B + ln] = ux[i];
    for (int i = 0; i < 36; ++i) a.ls.uP[i * B + ln] = uP[i];
    for (int i = 0; i < 6; ++i) xe[i] = ux[i];
  } else {
    for (int i = 0; i < 4; ++i) xe[i] = xn[i];
    xe[4] = xe[5] = 0.0;
  }
  const int variant = write_qp_params(a, ln, xe);
  a.ls.variant[ln] = variant;
  if (a.sc.in_track) {                      // in-place x/y swap of the stored estimate, simhelpers.py:72
    const double t = xe[0];
    xe[0] = xe[1];
    xe[1] = t;
  }
  for (int i = 0; i < 4; ++i) a.ls.xstore[i * B + ln] = xe[i];
  if (a.out.x_est)
    for (int i = 0; i < 6; ++i) a.out.x_est[((size_t)i * a.out.T1 + est_idx) * B + ln] = xe[i];
  return variant;
}

// ------------------------------------------------------------------------------------------
// Round 0: initial conditions (trajectorySimulate.py:248-269), first termination test, binning.
__global__ void init_kernel(const __grid_constant__ PostArgs a, const double *__restrict__ x0 /*[4][B]*/) {
  const int ln = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t B = a.B;
  bool live = false;
  int variant = 0;
  if (ln < a.B) {
    double x[6];
    for (int i = 0; i < 4; ++i) x[i] = x0[i * B + ln];
    x[4] = x[5] = 0.0;
    for (int i = 0; i < 4; ++i) {
      a.ls.xtrue[i * B + ln] = x[i];
      a.ls.xstore[i * B + ln] = x[i];
      a.ls.xfin[i * B + ln] = nan("");
    }
    for (int i = 0; i < 6; ++i) a.ls.ux[i * B + ln] = x[i];
    for (int i = 0; i < 36; ++i) a.ls.uP[i * B + ln] = (i % 7 == 0) ? ((i / 7 < 4) ? 1e-20 : 1.0) : 0.0;
    a.ls.uprev[ln] = a.ls.uprev[B + ln] = 0.0;
    a.ls.unext[ln] = a.ls.unext[B + ln] = 0.0;
    a.ls.xintf[ln] = 0.0;
    a.ls.noise[ln] = (a.sc.has_noise && a.noise_in) ? a.noise_in[ln] : 0.0;
    a.ls.noise[B + ln] = (a.sc.has_noise && a.noise_in) ? a.noise_in[B + ln] : 0.0;
    a.ls.step[ln] = 0;
    a.ls.succ[ln] = 0;
    a.ls.nsolve[ln] = 0;
    variant = write_qp_params(a, ln, x);
    a.ls.variant[ln] = variant;
    a.iter[ln] = 0;
    a.status[ln] = -10;
    const int T1 = a.out.T1;
    if (a.out.x_true) for (int i = 0; i < 4; ++i) a.out.x_true[((size_t)i * T1) * B + ln] = x[i];
    if (a.out.x_est) for (int i = 0; i < 6; ++i) a.out.x_est[((size_t)i * T1) * B + ln] = x[i];
    if (a.out.ctrl) for (int i = 0; i < 2; ++i) a.out.ctrl[((size_t)i * T1) * B + ln] = 0.0;
    const int first = (a.mode == MODE_CONTINUOUS) ? a.ratio : 0;
    a.ls.sub[ln] = first;
    if (a.mode == MODE_CONTINUOUS) {            // xtrueP[:, :ratio+1] = x0, ctrls[:, :ratio+1] = 0, seq[:ratio] = 0 (:289-292, :434)
      const size_t NS = a.out.NS;
      const int nfill = min(a.ratio + 1, a.n_sub_total);
      for (int s = 0; s < nfill; ++s) {
        if (a.out.x_true_sub) for (int i = 0; i < 4; ++i) a.out.x_true_sub[((size_t)i * NS + s) * B + ln] = x[i];
        if (a.out.ctrl_sub) for (int i = 0; i < 2; ++i) a.out.ctrl_sub[((size_t)i * NS + s) * B + ln] = 0.0;
      }
    }
    const bool nothing = (a.mode == MODE_CONTINUOUS) ? (first >= a.n_sub_total - 1) : (a.nsteps <= 0);
    if (nothing) {
      a.ls.iterm[ln] = (a.mode == MODE_CONTINUOUS) ? a.n_sub_total : a.nsteps;
      a.lane_state[ln] = LANE_FINISHED;
    } else if (terminated(a.sc, x)) {
      a.ls.iterm[ln] = first;
      a.lane_state[ln] = LANE_FINISHED;
    } else {
      a.ls.iterm[ln] = (a.mode == MODE_CONTINUOUS) ? a.n_sub_total : a.nsteps;
      a.lane_state[ln] = LANE_SOLVING;
      live = true;
    }
  }
  bin_lane(a, ln, live, variant, true);
}

// QP-only seam (mpcb_qp_solve): parameters from xhat, every lane solves.
__global__ void qp_prepare_kernel(const __grid_constant__ PostArgs a, const double *__restrict__ xhat /*[6][B]*/) {
  const int ln = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t B = a.B;
  bool live = false;
  int variant = 0;
  if (ln < a.B) {
    double xe[6];
    for (int i = 0; i < 6; ++i) xe[i] = xhat[i * B + ln];
    variant = write_qp_params(a, ln, xe);
    a.ls.variant[ln] = variant;
    a.iter[ln] = 0;
    a.status[ln] = -10;
    a.lane_state[ln] = LANE_SOLVING;
    live = true;
  }
  bin_lane(a, ln, live, variant);
}

// After every ADMM block: lanes whose solve finished run the rest of the control step
// (trajectorySimulate.py:298-356 / trajectorySimulateC.py:340-409), then all still-solving lanes
// are binned by sign variant for the next block.
__global__ void post_kernel(const __grid_constant__ PostArgs a) {
  const int ln = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t B = a.B;
  bool live = false, fresh = false;
  int variant = 0;
  if (blockIdx.x == 0 && threadIdx.x < 4) a.cnt_cur[threadIdx.x] = 0;   // consumed by the ADMM kernel before us
  if (ln < a.B) {
    uint8_t stt = a.lane_state[ln];
    variant = a.ls.variant[ln];
    if (stt == LANE_SOLVE_DONE && a.mode == MODE_QP_ONLY) {
      stt = LANE_FINISHED;
      a.lane_state[ln] = stt;
    } else if (stt == LANE_SOLVE_DONE) {
      const int T1 = a.out.T1;
      const int i = a.ls.step[ln];           // control step index (discrete) / solve index (continuous)
      const int status = a.status[ln];
      double xs[4], x[4], u[2], uraw[2], uprev[2], w[2], xn[4];
      for (int k = 0; k < 4; ++k) {
        xs[k] = a.ls.xstore[k * B + ln];
        x[k] = a.ls.xtrue[k * B + ln];
      }
      const int code = select_control(a, ln, status, xs, u, uraw);
      uprev[0] = a.ls.unext[ln];             // ctrls[:, i]: the command chosen at the previous step
      uprev[1] = a.ls.unext[B + ln];
      w[0] = a.ls.noise[ln];
      w[1] = a.ls.noise[B + ln];
      a.ls.nsolve[ln] += 1;
      if (a.out.status) a.out.status[(size_t)i * B + ln] = (int8_t)status;
      if (a.out.iters) a.out.iters[(size_t)i * B + ln] = (int16_t)a.iter[ln];
      if (a.out.rho_hist) a.out.rho_hist[(size_t)i * B + ln] = a.rho[ln];
      if (a.out.ctrlr_seq) a.out.ctrlr_seq[(size_t)i * B + ln] = (uint8_t)code;
      if (a.out.u_raw) {
        a.out.u_raw[((size_t)0 * (T1 - 1) + i) * B + ln] = uraw[0];
        a.out.u_raw[((size_t)1 * (T1 - 1) + i) * B + ln] = uraw[1];
      }
      if (a.out.ctrl) {
        a.out.ctrl[((size_t)0 * T1 + i + 1) * B + ln] = u[0];
        a.out.ctrl[((size_t)1 * T1 + i + 1) * B + ln] = u[1];
      }
      a.ls.unext[ln] = u[0];
      a.ls.unext[B + ln] = u[1];
      // success is scanned over x_true[:, 1 .. i_term-1] (:369-376); x is a live state here
      bool succ = a.ls.succ[ln] != 0;
      if ((i >= 1 || a.mode == MODE_CONTINUOUS) && success_cond(a.sc, x)) succ = true;
      double xfin[4] = {x[0], x[1], x[2], x[3]};

      bool fin = false;
      if (a.mode == MODE_DISCRETE) {
        plant_lin(a.sc, x, uprev, w, xn);
        variant = estimate_and_refresh(a, ln, xn, uprev, i + 1);
        if (a.out.x_true)
          for (int k = 0; k < 4; ++k) a.out.x_true[((size_t)k * T1 + i + 1) * B + ln] = xn[k];
        if (a.sc.has_noise && ((i + 1) % a.sc.noise_length == 0)) {
          const int r = min((i + 1) / a.sc.noise_length, a.n_refresh - 1);
          a.ls.noise[ln] = a.noise_in[((size_t)r * 2 + 0) * B + ln];
          a.ls.noise[B + ln] = a.noise_in[((size_t)r * 2 + 1) * B + ln];
        }
        a.ls.step[ln] = i + 1;
        if (i + 1 >= a.nsteps) {
          fin = true;                         // i_term stays nsim
        } else if (terminated(a.sc, xn)) {
          a.ls.iterm[ln] = i + 1;
          fin = true;
        }
      } else {
        // continuous (trajectorySimulateC.py:325-409): solve index i happens at substep (i+1)*ratio;
        // the plant sees the previous command on that substep, the new one from the next substep on.
        int sub = a.ls.sub[ln];
        const double nm = a.sc.mean_mtn, h = a.T_cont;
        for (int k = 0; k < 4; ++k) xn[k] = x[k];
        {
          const int nr = min(sub / a.noise_hold_sub, a.n_refresh - 1);
          const double w0 = a.noise_in ? a.noise_in[((size_t)nr * 2 + 0) * B + ln] : 0.0;
          const double w1 = a.noise_in ? a.noise_in[((size_t)nr * 2 + 1) * B + ln] : 0.0;
          if (!a.sc.delta_v) {
            rk4_substep(xn, uprev[0], uprev[1], nm, h);
          } else {
            rk4_substep(xn, 0.0, 0.0, nm, h);
            xn[2] += uprev[0];                // impulsive delta-v at the sample instant (:377-378)
            xn[3] += uprev[1];
          }
          xn[0] += w0;
          xn[1] += w1;
        }
        variant = estimate_and_refresh(a, ln, xn, uprev, i + 1);
        if (a.out.x_true)
          for (int k = 0; k < 4; ++k) a.out.x_true[((size_t)k * T1 + i + 1) * B + ln] = xn[k];
        const size_t NS = a.out.NS;
        auto put_sub = [&](int s) {               // substep s done: xtrueP[:, s+1], ctrls[:, s+1], seq[s]
          if (a.out.x_true_sub) for (int k = 0; k < 4; ++k) a.out.x_true_sub[((size_t)k * NS + s + 1) * B + ln] = xn[k];
          if (a.out.ctrl_sub) {
            a.out.ctrl_sub[((size_t)0 * NS + s + 1) * B + ln] = u[0];
            a.out.ctrl_sub[((size_t)1 * NS + s + 1) * B + ln] = u[1];
          }
          if (a.out.ctrlr_sub) a.out.ctrlr_sub[(size_t)s * B + ln] = (uint8_t)code;
        };
        put_sub(sub);
        sub += 1;
        const int next_sample = (i + 2) * a.ratio;
        const bool more_samples = (i + 2) < a.nsteps;       // disc_j < nsimD (:335)
        while (true) {
          if (sub >= a.n_sub_total - 1) {                   // range(ratio, nsimC-1) exhausted: i_term = nsimC
            if (success_cond(a.sc, xn)) succ = true;
            for (int k = 0; k < 4; ++k) xfin[k] = xn[k];
            fin = true;
            break;
          }
          if (terminated(a.sc, xn)) { a.ls.iterm[ln] = sub; fin = true; break; }
          if (more_samples && sub == next_sample) break;     // next solve (its epilogue scans this state)
          if (success_cond(a.sc, xn)) succ = true;
          for (int k = 0; k < 4; ++k) xfin[k] = xn[k];
          const int nr = min(sub / a.noise_hold_sub, a.n_refresh - 1);
          const double w0 = a.noise_in ? a.noise_in[((size_t)nr * 2 + 0) * B + ln] : 0.0;
          const double w1 = a.noise_in ? a.noise_in[((size_t)nr * 2 + 1) * B + ln] : 0.0;
          if (!a.sc.delta_v) rk4_substep(xn, u[0], u[1], nm, h);
          else rk4_substep(xn, 0.0, 0.0, nm, h);
          xn[0] += w0;
          xn[1] += w1;
          put_sub(sub);
          sub += 1;
        }
        a.ls.sub[ln] = sub;
        a.ls.step[ln] = i + 1;
      }
      a.ls.succ[ln] = succ ? 1 : 0;
      for (int k = 0; k < 4; ++k) a.ls.xfin[k * B + ln] = xfin[k];
      for (int k = 0; k < 4; ++k) a.ls.xtrue[k * B + ln] = xn[k];
      if (fin) {
        stt = LANE_FINISHED;
      } else {
        stt = LANE_SOLVING;
        a.iter[ln] = 0;
        a.status[ln] = -10;
        fresh = true;
      }
      a.lane_state[ln] = stt;
    }
    live = (stt == LANE_SOLVING);
  }
  bin_lane(a, ln, live, variant, fresh);
}

// Final per-lane results + batch statistics (test/disturbRejComp.py:87-100, success_rates_test.py:66-75).
__global__ void finalize_kernel(const __grid_constant__ PostArgs a, double *__restrict__ stats /*[10]*/,
                                const int *__restrict__ flip) {
  const int ln = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t B = a.B;
  double v[6] = {0, 0, 0, 0, 0, 0};
  if (ln < a.B) {
    double d2 = 0.0;
    for (int k = 0; k < 4; ++k) {
      const double d = a.ls.xfin[k * B + ln] - a.sc.xr[k];
      d2 += d * d;
    }
    const double fd = sqrt(d2);
    const int it = a.ls.iterm[ln];
    if (a.out.i_term) a.out.i_term[ln] = it;
    if (a.out.is_success) a.out.is_success[ln] = a.ls.succ[ln];
    if (a.out.final_dist) a.out.final_dist[ln] = fd;
    if (a.out.fd_all) a.out.fd_all[ln] = (fd == fd) ? fd : 0.0;      // stats[0], stats[1]: fixed-order sum afterwards
    v[2] = a.ls.succ[ln];
    v[3] = 1.0;
    v[4] = it;
    v[5] = a.ls.nsolve[ln];
    if (flip[ln]) atomicAdd(&stats[7], 1.0);
    if (a.out.ukf_clamped) a.out.ukf_clamped[ln] = a.ls.ukf_clamp[ln];
    if (a.ls.ukf_clamp[ln]) atomicAdd(&stats[8], 1.0);
    if (it < ((a.mode == MODE_CONTINUOUS) ? a.n_sub_total : a.nsteps)) atomicAdd(&stats[9], 1.0);
  }
#pragma unroll
  for (int k = 2; k < 6; ++k) {                  // integer counts: exact in any order
    const double s = warp_sum(v[k]);
    if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(&stats[k], s);
  }
}

// ---- unit seams -------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// Disturbance draws on the device: Philox4x32-10 (Salmon et al. SC'11 / Random123; oracle/philox_ref.py pins it to the
// published known-answer vectors).  One 4-word block per (lane, refresh): counter = (lane lo, lane hi, refresh, 0),
// key = seed; words 0,1 -> one Box-Muller pair = the two position disturbances, words 2,3 unused -- the reference draws
// random.normal(0,1,4) and uses two (src/trajectorySimulate.py:268, 351-356).
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
__global__ void noise_fill_kernel(int B, int n_refresh, double sx, double sy, unsigned long long seed, unsigned long long lane_offset,
                                  double *__restrict__ noise /*[R][2][B]*/, uint32_t *__restrict__ raw /*[R][4][B] or null*/) {
  const int ln = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (ln >= B || r >= n_refresh) return;
  const unsigned long long lane = (unsigned long long)ln + lane_offset;
  uint32_t c[4] = {(uint32_t)lane, (uint32_t)(lane >> 32), (uint32_t)r, 0u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  if (raw)
    for (int k = 0; k < 4; ++k) raw[((size_t)r * 4 + k) * B + ln] = c[k];
  const double u0 = ((double)c[0] + 0.5) * (1.0 / 4294967296.0), u1 = ((double)c[1] + 0.5) * (1.0 / 4294967296.0);
  const double rad = sqrt(-2.0 * log(u0));
  double sn, cs;
  sincospi(2.0 * u1, &sn, &cs);
  noise[((size_t)r * 2 + 0) * B + ln] = sx * rad * cs;
  noise[((size_t)r * 2 + 1) * B + ln] = sy * rad * sn;
}

__global__ void ukf_step_kernel(SimConst c, int B, double *x, double *P, const double *u, const double *z) {
  const int ln = blockIdx.x * blockDim.x + threadIdx.x;
  if (ln >= B) return;
  double ux[6], uP[36], uu[2] = {u[ln], u[(size_t)B + ln]}, zz[2] = {z[ln], z[(size_t)B + ln]};
  for (int i = 0; i < 6; ++i) ux[i] = x[(size_t)i * B + ln];
  for (int i = 0; i < 36; ++i) uP[i] = P[(size_t)i * B + ln];
  if (c.estimator == MPCB_EST_KF) kf_step(c, ux, uP, uu, zz);      // z = measured position
  else ukf_step(c, ux, uP, uu, zz);
  for (int i = 0; i < 6; ++i) x[(size_t)i * B + ln] = ux[i];
  for (int i = 0; i < 36; ++i) P[(size_t)i * B + ln] = uP[i];
}

__global__ void plant_lin_kernel(SimConst c, int B, double *x, const double *u, const double *w) {
  const int ln = blockIdx.x * blockDim.x + threadIdx.x;
  if (ln >= B) return;
  double xx[4], xn[4], uu[2] = {u[ln], u[(size_t)B + ln]}, ww[2] = {w ? w[ln] : 0.0, w ? w[(size_t)B + ln] : 0.0};
  for (int i = 0; i < 4; ++i) xx[i] = x[(size_t)i * B + ln];
  plant_lin(c, xx, uu, ww, xn);
  for (int i = 0; i < 4; ++i) x[(size_t)i * B + ln] = xn[i];
}

__global__ void plant_rk4_kernel(SimConst c, int B, double *x, const double *u, const double *w, int nsub, double dt) {
  const int ln = blockIdx.x * blockDim.x + threadIdx.x;
  if (ln >= B) return;
  double xx[4];
  const double u0 = u[ln], u1 = u[(size_t)B + ln], w0 = w ? w[ln] : 0.0, w1 = w ? w[(size_t)B + ln] : 0.0;
  for (int i = 0; i < 4; ++i) xx[i] = x[(size_t)i * B + ln];
  for (int s = 0; s < nsub; ++s) {
    rk4_substep(xx, u0, u1, c.mean_mtn, dt);
    xx[0] += w0;
    xx[1] += w1;
  }
  for (int i = 0; i < 4; ++i) x[(size_t)i * B + ln] = xx[i];
}
