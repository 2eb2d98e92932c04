/*
 * mpc_ref.c -- ORACLE (test infrastructure, never linked into the product): a compiled CPU twin of the reference's
 * per-trajectory closed loop, one trajectory at a time, plain C, float64.
 *
 * What it restates, line for line with oracle/batched_ref.py / oracle/sim_ref.py (which carry the file:line citations
 * into the reference):
 *   src/trajectorySimulate.py:285-356   the control-step loop: solve -> controller select (:299-314) -> sequential norm clip
 *                                       (:317-319) -> plant step with one-step actuation delay (:323-324) -> UKF (:329-337)
 *                                       -> bound / sign update (:340-348) -> noise hold (:351-356); termination (:288-293),
 *                                       success scan (:369-376)
 *   src/simhelpers.py:66-67,116-138     velocity signs, the bounds that move with the estimate
 *   osqp 0.6.x (absent from /root/reference: PARITY UNPINNED, see oracle/__init__.py)   ADMM in the Ruiz-scaled variables,
 *                                       rho_vec with the 1e3 equality factor, alpha = 1.6, check_termination every 25
 *                                       iterations on unscaled residuals, primal-infeasibility certificate, adaptive rho on
 *                                       scaled residuals every adaptive_rho_interval iterations, warm start across steps
 *   filterpy 1.4.5 UnscentedKalmanFilter (unpinned likewise)   Merwe points n = 6, alpha = .1, beta = 2, kappa = -1, R = 0,
 *                                       points regenerated after predict; a non-positive Cholesky pivot is clamped
 *
 * Linear algebra: OSQP solves the quasi-definite KKT system with a sparse LDL'; eliminating the constraint block gives
 * (P + sigma I + A' diag(rho_vec) A) x~ = sigma x - q + A'(rho_vec z - y), z~ = A x~ -- the same iterates -- which is
 * factored here with a dense Cholesky (n <= 201) whenever the lane's rho or velocity-sign variant changes.  A is kept as
 * CSR.  This is the CPU baseline of bench.py ("port": osqp / filterpy / control cannot be installed offline) and a
 * second, independent implementation the GPU results are compared with.
 *
 * Input: the host tables of include/mpcb.h (struct mpcb_problem, filled by mpc_arpo_project_b200/engine.py's code path
 * from problem.py) and the same SoA batch layout as the C ABI, so one fixture drives both.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "../../include/mpcb.h"

#define RHO_MIN 1e-6
#define RHO_MAX 1e6
#define RHO_EQ 1e3
#define RHO_TOL 1e-4
#define DIV_TOL 1e-30
#define UKF_NPL 0.05
#define UKF_WM0 (-5.95 / 0.05)
#define UKF_WC0 (-5.95 / 0.05 + (1.0 - 0.01 + 2.0))
#define UKF_WI (0.5 / 0.05)

typedef struct {
  int n, m, nnz;
  int *rp, *ci;          /* CSR of A (pattern shared by the four sign variants) */
  double *val;           /* variant 0 values */
  int *sgn;              /* per entry: 0 plain, 1 carries C1 (sign of vx^), 2 carries C2 */
  double *Pb;            /* P_s + sigma I, dense n x n */
  uint8_t *inf_l, *inf_u;
} shared_t;

typedef struct {
  double *x, *z, *y, *l, *u, *dy, *rv, *rinv, *av, *M, *rhs, *xt, *zt, *Ax, *Px, *Aty, *t1;
  double rho, f_rho;
  int f_var;             /* (rho, variant, re-typed rows) the factor in M was built for */
  unsigned eqmask, f_eqmask;   /* bit k: the velocity-bound row of stage k is currently an EQUALITY for OSQP (u - l < RHO_TOL after
                                  scaling, osqp auxil.c update_rho_vec: rho_vec[i] = 1e3 rho and the KKT matrix is refactored) */
} work_t;

static void csr_mv(const shared_t *S, const double *av, const double *x, double *out) {
  for (int i = 0; i < S->m; ++i) {
    double acc = 0.0;
    for (int e = S->rp[i]; e < S->rp[i + 1]; ++e) acc += av[e] * x[S->ci[e]];
    out[i] = acc;
  }
}
static void csr_mtv(const shared_t *S, const double *av, const double *v, double *out) {
  for (int j = 0; j < S->n; ++j) out[j] = 0.0;
  for (int i = 0; i < S->m; ++i)
    for (int e = S->rp[i]; e < S->rp[i + 1]; ++e) out[S->ci[e]] += av[e] * v[i];
}

/* M = Pb + A' diag(rv) A, lower Cholesky in place */
static int factor(const shared_t *S, work_t *w) {
  const int n = S->n;
  memcpy(w->M, S->Pb, sizeof(double) * n * n);
  for (int i = 0; i < S->m; ++i)
    for (int e = S->rp[i]; e < S->rp[i + 1]; ++e) {
      const double a = w->rv[i] * w->av[e];
      const int r = S->ci[e];
      for (int f = S->rp[i]; f < S->rp[i + 1]; ++f) w->M[(size_t)r * n + S->ci[f]] += a * w->av[f];
    }
  for (int j = 0; j < n; ++j) {
    double d = w->M[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= w->M[(size_t)j * n + k] * w->M[(size_t)j * n + k];
    if (!(d > 0.0)) return -1;
    d = sqrt(d);
    w->M[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double v = w->M[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) v -= w->M[(size_t)i * n + k] * w->M[(size_t)j * n + k];
      w->M[(size_t)i * n + j] = v / d;
    }
  }
  return 0;
}
static void chol_solve(const double *L, int n, double *b) {
  for (int i = 0; i < n; ++i) {
    double v = b[i];
    for (int k = 0; k < i; ++k) v -= L[(size_t)i * n + k] * b[k];
    b[i] = v / L[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double v = b[i];
    for (int k = i + 1; k < n; ++k) v -= L[(size_t)k * n + i] * b[k];
    b[i] = v / L[(size_t)i * n + i];
  }
}

static void set_rho_vec(const mpcb_problem *p, work_t *w) {
  const int nX = 4 * (p->Nx + 1);
  for (int i = 0; i < p->m; ++i) {
    w->rv[i] = p->ctype[i] == -1 ? RHO_MIN : (p->ctype[i] == 1 ? RHO_EQ * w->rho : w->rho);
    w->rinv[i] = 1.0 / w->rv[i];
  }
  for (int k = 0; k <= p->Nb && k < 32; ++k)
    if (w->eqmask & (1u << k)) {
      w->rv[nX + 5 * k + 3] = RHO_EQ * w->rho;
      w->rinv[nX + 5 * k + 3] = 1.0 / w->rv[nX + 5 * k + 3];
    }
}

/* prob.solve(): returns OSQP status_val, *iters_out = info.iter */
static int solve(const mpcb_problem *p, const shared_t *S, work_t *w, int variant, int *iters_out, int64_t *iter_total) {
  const int n = p->n, m = p->m;
  const double sigma = p->sigma, alpha = p->alpha, cinv = 1.0 / p->c;
  for (int e = 0; e < S->nnz; ++e) {
    const int sg = S->sgn[e];
    w->av[e] = ((sg == 1 && (variant & 1)) || (sg == 2 && (variant & 2))) ? -S->val[e] : S->val[e];
  }
  double qn_u = 0.0, qn_s = 0.0;
  for (int j = 0; j < n; ++j) {
    qn_u = fmax(qn_u, fabs(p->q_s[j] / p->D[j]));
    qn_s = fmax(qn_s, fabs(p->q_s[j]));
  }
  int it = 0, status = MPCB_QP_UNSOLVED;
  set_rho_vec(p, w);
  while (it < p->max_iter) {
    if (w->f_rho != w->rho || w->f_var != variant || w->f_eqmask != w->eqmask) {
      set_rho_vec(p, w);
      if (factor(S, w) != 0) return MPCB_QP_UNSOLVED;
      w->f_rho = w->rho;
      w->f_var = variant;
      w->f_eqmask = w->eqmask;
    }
    for (int k = 0; k < p->check_termination; ++k) {
      for (int i = 0; i < m; ++i) w->t1[i] = w->rv[i] * w->z[i] - w->y[i];
      csr_mtv(S, w->av, w->t1, w->rhs);
      for (int j = 0; j < n; ++j) w->xt[j] = sigma * w->x[j] - p->q_s[j] + w->rhs[j];
      chol_solve(w->M, n, w->xt);
      csr_mv(S, w->av, w->xt, w->zt);
      for (int j = 0; j < n; ++j) w->x[j] = alpha * w->xt[j] + (1.0 - alpha) * w->x[j];
      for (int i = 0; i < m; ++i) {
        const double zr = alpha * w->zt[i] + (1.0 - alpha) * w->z[i];
        const double zn = fmin(fmax(zr + w->rinv[i] * w->y[i], w->l[i]), w->u[i]);
        w->dy[i] = w->rv[i] * (zr - zn);
        w->y[i] += w->dy[i];
        w->z[i] = zn;
      }
    }
    it += p->check_termination;
    *iter_total += p->check_termination;
    /* update_info */
    csr_mv(S, w->av, w->x, w->Ax);
    for (int i = 0; i < n; ++i) {
      double acc = 0.0;
      for (int j = 0; j < n; ++j) acc += (S->Pb[(size_t)i * n + j] - (i == j ? sigma : 0.0)) * w->x[j];
      w->Px[i] = acc;
    }
    csr_mtv(S, w->av, w->y, w->Aty);
    double pri_u = 0, nz_u = 0, nax_u = 0, pri_s = 0, nz_s = 0, nax_s = 0;
    for (int i = 0; i < m; ++i) {
      const double ei = 1.0 / p->E[i], pv = w->Ax[i] - w->z[i];
      pri_u = fmax(pri_u, fabs(ei * pv)); nz_u = fmax(nz_u, fabs(ei * w->z[i])); nax_u = fmax(nax_u, fabs(ei * w->Ax[i]));
      pri_s = fmax(pri_s, fabs(pv)); nz_s = fmax(nz_s, fabs(w->z[i])); nax_s = fmax(nax_s, fabs(w->Ax[i]));
    }
    double dua_u = 0, npx_u = 0, naty_u = 0, dua_s = 0, npx_s = 0, naty_s = 0;
    for (int j = 0; j < n; ++j) {
      const double di = 1.0 / p->D[j], dv = p->q_s[j] + w->Px[j] + w->Aty[j];
      dua_u = fmax(dua_u, fabs(di * dv)); npx_u = fmax(npx_u, fabs(di * w->Px[j])); naty_u = fmax(naty_u, fabs(di * w->Aty[j]));
      dua_s = fmax(dua_s, fabs(dv)); npx_s = fmax(npx_s, fabs(w->Px[j])); naty_s = fmax(naty_s, fabs(w->Aty[j]));
    }
    dua_u *= cinv;
    /* primal infeasibility certificate on the last delta_y */
    double ndy = 0.0, lhs = 0.0;
    for (int i = 0; i < m; ++i) {
      double d = w->dy[i];
      if (S->inf_l[i] && S->inf_u[i]) d = 0.0;
      else if (S->inf_u[i]) d = fmin(d, 0.0);
      else if (S->inf_l[i]) d = fmax(d, 0.0);
      w->t1[i] = d;
      ndy = fmax(ndy, fabs(p->E[i] * d));
      lhs += w->u[i] * fmax(d, 0.0) + w->l[i] * fmin(d, 0.0);
    }
    csr_mtv(S, w->av, w->t1, w->rhs);
    double natdy = 0.0;
    for (int j = 0; j < n; ++j) natdy = fmax(natdy, fabs(w->rhs[j] / p->D[j]));
    for (int pass = 0; pass < 2; ++pass) {
      const double k = pass ? 10.0 : 1.0;
      if (pass && it < p->max_iter) break;
      const double eps_p = k * p->eps_abs + k * p->eps_rel * fmax(nz_u, nax_u);
      const double eps_d = k * p->eps_abs + k * p->eps_rel * cinv * fmax(qn_u, fmax(naty_u, npx_u));
      const int prim_ok = pri_u < eps_p, dual_ok = dua_u < eps_d;
      if (prim_ok && dual_ok) { status = pass ? MPCB_QP_SOLVED_INACCURATE : MPCB_QP_SOLVED; break; }
      if (!prim_ok) {
        const double eps_i = k * p->eps_prim_inf;
        if (ndy > DIV_TOL && lhs < -eps_i * ndy && natdy < eps_i * ndy) {
          status = pass ? MPCB_QP_PRIMAL_INFEASIBLE_INACCURATE : MPCB_QP_PRIMAL_INFEASIBLE;
          break;
        }
      }
      if (pass == 0 && p->adaptive_rho && (it % p->adaptive_rho_interval == 0)) {
        const double pr = pri_s / (fmax(nz_s, nax_s) + 1e-10);
        const double du = dua_s / (fmax(qn_s, fmax(naty_s, npx_s)) + 1e-10);
        double est = w->rho * sqrt(pr / (du + 1e-10));
        est = fmin(fmax(est, RHO_MIN), RHO_MAX);
        if (est > w->rho * p->adaptive_rho_tolerance || est < w->rho / p->adaptive_rho_tolerance) w->rho = est;
      }
    }
    if (status != MPCB_QP_UNSOLVED) break;
    if (it >= p->max_iter) { status = MPCB_QP_MAX_ITER; break; }
  }
  *iters_out = it;
  return status;
}

/* upper Cholesky U'U = s P with non-positive pivots clamped (row stays zero); returns 0 if a pivot was clamped */
static int chol_upper6(const double *P, double s, double *U) {
  int ok = 1;
  memset(U, 0, sizeof(double) * 36);
  for (int i = 0; i < 6; ++i) {
    double d = s * P[i * 6 + i];
    for (int k = 0; k < i; ++k) d -= U[k * 6 + i] * U[k * 6 + i];
    if (!(d > 0.0)) { ok = 0; continue; }
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
static int sigma_points(const double *x, const double *P, double *sig) {
  double U[36];
  const int ok = chol_upper6(P, UKF_NPL, U);
  for (int j = 0; j < 6; ++j) sig[j] = x[j];
  for (int k = 0; k < 6; ++k)
    for (int j = 0; j < 6; ++j) {
      sig[(k + 1) * 6 + j] = x[j] + U[k * 6 + j];
      sig[(k + 7) * 6 + j] = x[j] - U[k * 6 + j];
    }
  return ok;
}
/* kf.predict(u); kf.update(z), R = 0 */
static int ukf_step(const mpcb_problem *p, double *x, double *P, const double *u, const double *zm) {
  double sig[78], sf[78], xm[6], Pm[36];
  int ok = sigma_points(x, P, sig);
  for (int k = 0; k < 13; ++k)
    for (int i = 0; i < 6; ++i) {
      double acc = 0.0;
      for (int j = 0; j < 6; ++j) acc += p->Ao[i * 6 + j] * sig[k * 6 + j];
      sf[k * 6 + i] = acc + p->Bou[i * 2] * u[0] + p->Bou[i * 2 + 1] * u[1];
    }
  for (int i = 0; i < 6; ++i) {
    double acc = UKF_WM0 * sf[i];
    for (int k = 1; k < 13; ++k) acc += UKF_WI * sf[k * 6 + i];
    xm[i] = acc;
  }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double acc = 0.0;
      for (int k = 0; k < 13; ++k) acc += (k == 0 ? UKF_WC0 : UKF_WI) * (sf[k * 6 + i] - xm[i]) * (sf[k * 6 + j] - xm[j]);
      Pm[i * 6 + j] = acc + p->Qw[i * 6 + j];
    }
  ok &= sigma_points(xm, Pm, sf);
  double zs[26], zp[2] = {0.0, 0.0};
  for (int k = 0; k < 13; ++k) {
    const double a = sf[k * 6], b = sf[k * 6 + 1];
    zs[k * 2] = sqrt(a * a + b * b);
    zs[k * 2 + 1] = atan2(b, a);
    const double wk = (k == 0 ? UKF_WM0 : UKF_WI);
    zp[0] += wk * zs[k * 2];
    zp[1] += wk * zs[k * 2 + 1];
  }
  double Sm[4] = {0, 0, 0, 0}, Pxz[12];
  memset(Pxz, 0, sizeof Pxz);
  for (int k = 0; k < 13; ++k) {
    const double wk = (k == 0 ? UKF_WC0 : UKF_WI);
    const double d0 = zs[k * 2] - zp[0], d1 = zs[k * 2 + 1] - zp[1];
    Sm[0] += wk * d0 * d0; Sm[1] += wk * d0 * d1; Sm[2] += wk * d1 * d0; Sm[3] += wk * d1 * d1;
    for (int i = 0; i < 6; ++i) {
      const double dx = sf[k * 6 + i] - xm[i];
      Pxz[i * 2] += wk * dx * d0;
      Pxz[i * 2 + 1] += wk * dx * d1;
    }
  }
  const double det = Sm[0] * Sm[3] - Sm[1] * Sm[2];
  const double SI[4] = {Sm[3] / det, -Sm[1] / det, -Sm[2] / det, Sm[0] / det};
  double K[12];
  for (int i = 0; i < 6; ++i) {
    K[i * 2] = Pxz[i * 2] * SI[0] + Pxz[i * 2 + 1] * SI[2];
    K[i * 2 + 1] = Pxz[i * 2] * SI[1] + Pxz[i * 2 + 1] * SI[3];
  }
  const double y0 = zm[0] - zp[0], y1 = zm[1] - zp[1];
  for (int i = 0; i < 6; ++i) x[i] = xm[i] + K[i * 2] * y0 + K[i * 2 + 1] * y1;
  for (int i = 0; i < 6; ++i) {
    const double ks0 = K[i * 2] * Sm[0] + K[i * 2 + 1] * Sm[2], ks1 = K[i * 2] * Sm[1] + K[i * 2 + 1] * Sm[3];
    for (int j = 0; j < 6; ++j) P[i * 6 + j] = Pm[i * 6 + j] - (ks0 * K[j * 2] + ks1 * K[j * 2 + 1]);
  }
  return ok;
}

static int terminated(const mpcb_problem *p, const double *x) {
  const double r = sqrt(x[0] * x[0] + x[1] * x[1]);
  return (r < p->r_p) || ((p->in_track ? x[1] : x[0]) < p->r_p - p->r_tol);
}
static int success_cond(const mpcb_problem *p, const double *x) {
  const double dx = x[0] - p->xr[0], dy = x[1] - p->xr[1];
  if (!(sqrt(dx * dx + dy * dy) <= p->suc_dist)) return 0;
  return fabs(atan(x[3] / x[2])) * (180.0 / 3.141592653589793) <= p->suc_ang_deg;
}

/* the two prob.update calls (:340-348): bounds that move with the estimate; returns the velocity-sign variant */
static int set_params(const mpcb_problem *p, work_t *w, const double *xe) {
  const int m = p->m, nX = 4 * (p->Nx + 1);
  for (int i = 0; i < 4; ++i) w->l[i] = w->u[i] = -xe[i] * p->E[i];
  const double val = fabs(xe[0] - p->xr[0]) + fabs(xe[1] - p->xr[1]);
  w->eqmask = 0;
  for (int k = 0; k <= p->Nb; ++k) {
    const int r = nX + 5 * k + 3;
    w->u[r] = val * p->E[r];
    if (k < 32 && w->u[r] - w->l[r] < RHO_TOL) w->eqmask |= 1u << k;          /* row re-typing */
  }
  for (int i = 0; i < 2; ++i) w->l[m - 2 + i] = w->u[m - 2 + i] = (p->is_reject ? xe[4 + i] : 0.0) * p->E[m - 2 + i];
  return (xe[2] >= 0 ? 0 : 1) + (xe[3] >= 0 ? 0 : 2);
}

static void *zalloc(size_t n) { return calloc(n ? n : 1, 1); }

typedef struct {
  const mpcb_problem *p;
  const shared_t *S;
  int64_t B;
  int32_t nsteps, n_refresh;
  const double *x0, *noise;
  const mpcb_sim_out *out;
  int64_t *next;                 /* shared lane counter */
  int64_t solves, iters;         /* this thread's totals */
} job_t;

static void *worker(void *arg) {
  job_t *J = (job_t *)arg;
  const mpcb_problem *p = J->p;
  const shared_t SS = *J->S;
#define S SS
  const int64_t B = J->B;
  const int32_t nsteps = J->nsteps, n_refresh = J->n_refresh;
  const double *x0 = J->x0, *noise = J->noise;
  const mpcb_sim_out *out = J->out;
  const int n = p->n, m = p->m, nX = 4 * (p->Nx + 1);
  const size_t T1 = (size_t)nsteps + 1;
  const int nl = p->noise_length > 0 ? p->noise_length : 1;
  int64_t tot_solves = 0, tot_iters = 0;
  {
    work_t w;
    memset(&w, 0, sizeof w);
    w.x = (double *)zalloc(sizeof(double) * n); w.xt = (double *)zalloc(sizeof(double) * n); w.rhs = (double *)zalloc(sizeof(double) * n);
    w.Px = (double *)zalloc(sizeof(double) * n); w.Aty = (double *)zalloc(sizeof(double) * n);
    w.z = (double *)zalloc(sizeof(double) * m); w.y = (double *)zalloc(sizeof(double) * m); w.l = (double *)zalloc(sizeof(double) * m);
    w.u = (double *)zalloc(sizeof(double) * m); w.dy = (double *)zalloc(sizeof(double) * m); w.rv = (double *)zalloc(sizeof(double) * m);
    w.rinv = (double *)zalloc(sizeof(double) * m); w.zt = (double *)zalloc(sizeof(double) * m); w.Ax = (double *)zalloc(sizeof(double) * m);
    w.t1 = (double *)zalloc(sizeof(double) * m);
    w.av = (double *)zalloc(sizeof(double) * S.nnz);
    w.M = (double *)zalloc(sizeof(double) * n * n);
    for (;;) {
      const int64_t b = __atomic_fetch_add(J->next, 1, __ATOMIC_RELAXED);
      if (b >= B) break;
      /* osqp.setup: cold start */
      memset(w.x, 0, sizeof(double) * n); memset(w.z, 0, sizeof(double) * m); memset(w.y, 0, sizeof(double) * m);
      memset(w.dy, 0, sizeof(double) * m);
      memcpy(w.l, p->l_s, sizeof(double) * m);
      memcpy(w.u, p->u_s, sizeof(double) * m);
      w.rho = fmin(fmax(p->rho0, RHO_MIN), RHO_MAX);
      w.f_rho = -1.0;
      w.f_var = -1;
      w.eqmask = w.f_eqmask = 0;
      double xt[4], xe[6], ux[6], uP[36], xstore[4], unext[2] = {0, 0}, nz[2] = {0, 0}, xfin[4], xintf = 0.0;
      for (int k = 0; k < 4; ++k) { xt[k] = x0[(size_t)k * B + b]; xe[k] = xt[k]; xstore[k] = xt[k]; xfin[k] = NAN; }
      xe[4] = xe[5] = 0.0;
      memcpy(ux, xe, sizeof ux);
      for (int k = 0; k < 36; ++k) uP[k] = (k % 7 == 0) ? ((k / 7 < 4) ? 1e-20 : 1.0) : 0.0;
      if (p->has_noise) { nz[0] = noise[(size_t)0 * B + b]; nz[1] = noise[(size_t)1 * B + b]; }
      int variant = set_params(p, &w, xe);
      if (out->x_true) for (int k = 0; k < 4; ++k) out->x_true[((size_t)k * T1) * B + b] = xt[k];
      if (out->x_est) for (int k = 0; k < 6; ++k) out->x_est[((size_t)k * T1) * B + b] = xe[k];
      if (out->ctrl) for (int k = 0; k < 2; ++k) out->ctrl[((size_t)k * T1) * B + b] = 0.0;
      int iterm = nsteps, succ = 0, clamped = 0;
      for (int i = 0; i < nsteps; ++i) {
        if (terminated(p, xt)) { iterm = i; break; }
        int its = 0;
        const int st = solve(p, &S, &w, variant, &its, &tot_iters);
        tot_solves += 1;
        double u[2], uraw[2];
        int code;
        if (st != MPCB_QP_SOLVED) {
          xintf = xintf + xstore[0] - p->xr[0];
          for (int r = 0; r < 2; ++r) {
            double acc = 0.0;
            for (int j = 0; j < 4; ++j) acc += p->Kpf[r * 4 + j] * xstore[j];
            u[r] = -acc - p->Kif[r] * xintf;
          }
          code = MPCB_CTRL_FAILSAFE;
        } else {
          xintf = 0.0;
          u[0] = p->D[nX] * w.x[nX];
          u[1] = p->D[nX + 1] * w.x[nX + 1];
          code = MPCB_CTRL_MPC;
        }
        uraw[0] = u[0]; uraw[1] = u[1];
        const double nrm = sqrt(u[0] * u[0] + u[1] * u[1]);
        if (nrm > p->umax0) {
          u[0] = u[0] * (p->umax0 / nrm);
          const double nrm2 = sqrt(u[0] * u[0] + u[1] * u[1]);
          u[1] = u[1] * (p->umax0 / nrm2);
        }
        if (out->status) out->status[(size_t)i * B + b] = (int8_t)st;
        if (out->iters) out->iters[(size_t)i * B + b] = (int16_t)its;
        if (out->rho) out->rho[(size_t)i * B + b] = w.rho;
        if (out->ctrlr_seq) out->ctrlr_seq[(size_t)i * B + b] = (uint8_t)code;
        if (out->u_raw) { out->u_raw[((size_t)0 * (T1 - 1) + i) * B + b] = uraw[0]; out->u_raw[((size_t)1 * (T1 - 1) + i) * B + b] = uraw[1]; }
        if (out->ctrl) { out->ctrl[((size_t)0 * T1 + i + 1) * B + b] = u[0]; out->ctrl[((size_t)1 * T1 + i + 1) * B + b] = u[1]; }
        if (i >= 1 && success_cond(p, xt)) succ = 1;
        memcpy(xfin, xt, sizeof xfin);
        /* plant: x+ = Ad x + Bd u_prev + w  (one-step actuation delay) */
        double xn[4];
        for (int r = 0; r < 4; ++r) {
          double acc = 0.0;
          for (int j = 0; j < 4; ++j) acc += p->Ad[r * 4 + j] * xt[j];
          acc += p->Bd[r * 2] * unext[0] + p->Bd[r * 2 + 1] * unext[1];
          xn[r] = acc + (r < 2 ? nz[r] : 0.0);
        }
        if (p->has_noise) {
          const double zm[2] = {sqrt(xn[0] * xn[0] + xn[1] * xn[1]), atan2(xn[1], xn[0])};
          if (!ukf_step(p, ux, uP, unext, zm)) clamped = 1;
          memcpy(xe, ux, sizeof xe);
        } else {
          memcpy(xe, xn, sizeof(double) * 4);
          xe[4] = xe[5] = 0.0;
        }
        unext[0] = u[0]; unext[1] = u[1];
        variant = set_params(p, &w, xe);
        if (p->in_track) { const double t = xe[0]; xe[0] = xe[1]; xe[1] = t; }
        memcpy(xstore, xe, sizeof xstore);
        if (out->x_est) for (int k = 0; k < 6; ++k) out->x_est[((size_t)k * T1 + i + 1) * B + b] = xe[k];
        if (out->x_true) for (int k = 0; k < 4; ++k) out->x_true[((size_t)k * T1 + i + 1) * B + b] = xn[k];
        memcpy(xt, xn, sizeof xt);
        if (p->has_noise && ((i + 1) % nl == 0)) {
          const int r = (i + 1) / nl < n_refresh ? (i + 1) / nl : n_refresh - 1;
          nz[0] = noise[((size_t)r * 2 + 0) * B + b];
          nz[1] = noise[((size_t)r * 2 + 1) * B + b];
        }
      }
      double d2 = 0.0;
      for (int k = 0; k < 4; ++k) d2 += (xfin[k] - p->xr[k]) * (xfin[k] - p->xr[k]);
      if (out->i_term) out->i_term[b] = iterm;
      if (out->is_success) out->is_success[b] = succ;
      if (out->final_dist) out->final_dist[b] = sqrt(d2);
      if (out->ukf_clamped) out->ukf_clamped[b] = clamped;
    }
    free(w.x); free(w.xt); free(w.rhs); free(w.Px); free(w.Aty); free(w.z); free(w.y); free(w.l); free(w.u); free(w.dy);
    free(w.rv); free(w.rinv); free(w.zt); free(w.Ax); free(w.t1); free(w.av); free(w.M);
  }
#undef S
  J->solves = tot_solves;
  J->iters = tot_iters;
  return NULL;
}

int mpcref_abi_version(void) { return MPCB_ABI_VERSION; }

/* trajectorySimulate for B lanes, one lane at a time per thread.  Same argument meaning and array layouts as
 * mpcb_simulate_discrete (host pointers).  counts[0] = QP solves, counts[1] = ADMM iterations. */
int mpcref_simulate_discrete(const mpcb_problem *p, int64_t B, int32_t nsteps, const double *x0, const double *noise,
                             int32_t n_refresh, const mpcb_sim_out *out, int nthreads, int64_t *counts) {
  if (!p || !x0 || !out || B < 1 || nsteps < 0 || p->has_debris || p->estimator != MPCB_EST_UKF) return MPCB_ERR_INVALID;
  if (p->has_noise && (!noise || n_refresh < nsteps / (p->noise_length > 0 ? p->noise_length : 1) + 1)) return MPCB_ERR_INVALID;
  const int n = p->n, m = p->m, nX = 4 * (p->Nx + 1);
  shared_t S;
  memset(&S, 0, sizeof S);
  S.n = n; S.m = m;
  S.rp = (int *)zalloc(sizeof(int) * (m + 1));
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) S.nnz += p->A_s[(size_t)i * n + j] != 0.0;
  S.ci = (int *)zalloc(sizeof(int) * S.nnz);
  S.val = (double *)zalloc(sizeof(double) * S.nnz);
  S.sgn = (int *)zalloc(sizeof(int) * S.nnz);
  int e = 0;
  for (int i = 0; i < m; ++i) {
    S.rp[i] = e;
    for (int j = 0; j < n; ++j) {
      const double v = p->A_s[(size_t)i * n + j];
      if (v != 0.0) {
        S.ci[e] = j;
        S.val[e] = v;
        const int k = (i - nX) / 5;
        if (i >= nX && i < nX + 5 * (p->Nx + 1) && (i - nX) % 5 == 3) S.sgn[e] = (j == 4 * k + 2) ? 1 : ((j == 4 * k + 3) ? 2 : 0);
        ++e;
      }
    }
  }
  S.rp[m] = e;
  S.Pb = (double *)zalloc(sizeof(double) * n * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) S.Pb[(size_t)i * n + j] = p->P_s[(size_t)i * n + j] + (i == j ? p->sigma : 0.0);
  S.inf_l = (uint8_t *)zalloc(m);
  S.inf_u = (uint8_t *)zalloc(m);
  for (int i = 0; i < m; ++i) {
    S.inf_l[i] = p->l_s[i] < -1e30 * 1e-4;
    S.inf_u[i] = p->u_s[i] > 1e30 * 1e-4;
  }
  int64_t next = 0, tot_solves = 0, tot_iters = 0;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if ((int64_t)nthreads > B) nthreads = (int)B;
  job_t jobs[256];
  pthread_t th[256];
  for (int t = 0; t < nthreads; ++t) {
    job_t j = {p, &S, B, nsteps, n_refresh, x0, noise, out, &next, 0, 0};
    jobs[t] = j;
  }
  for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, worker, &jobs[t]);
  worker(&jobs[0]);
  for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
  for (int t = 0; t < nthreads; ++t) { tot_solves += jobs[t].solves; tot_iters += jobs[t].iters; }
  free(S.rp); free(S.ci); free(S.val); free(S.sgn); free(S.Pb); free(S.inf_l); free(S.inf_u);
  if (counts) { counts[0] = tot_solves; counts[1] = tot_iters; }
  return MPCB_OK;
}
