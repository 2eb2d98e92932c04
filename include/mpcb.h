/*
 * mpcb.h -- C ABI of libmpcb.so, the B200 (sm_100a) batched closed-loop MPC engine.
 *
 * The reference (IsaacTroche1/MPC_ARPO_Project) is pure Python and has no FFI; the seam this
 * library replaces is the per-control-step loop of
 *     src/trajectorySimulate.py:285-356   and   src/trajectorySimulateC.py:325-410
 * and, inside it, the OSQP object protocol (setup / solve / update) of
 *     src/trajectorySimulate.py:242-245, :296, :340-348.
 * Each entry point names the reference lines it stands in for.  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative mpcb_status on failure and never throws;
 *     mpcb_last_error() returns a thread-local message for the last failure;
 *   - one handle = one problem family (fixed horizons, weights, plant) on one GPU; calls on a
 *     handle are serialised on its CUDA stream;
 *   - batch arrays are double precision, structure-of-arrays "[field][B]" (lane index fastest);
 *   - `io_on_device` != 0: the batch pointers are device pointers on the handle's GPU;
 *     `io_on_device` == 0: they are host pointers and the call copies in/out itself;
 *   - stream contract: the handle works on its own stream (mpcb_stream).  Every call returns after that stream
 *     has drained, so outputs are complete on return.  INPUTS passed as device pointers must be complete before
 *     the call: either synchronise the producing stream, or call mpcb_wait_stream(h, producer) first, which makes the
 *     handle's stream wait (on the device, no host block) for everything enqueued on `producer` so far.
 */
#ifndef MPCB_H
#define MPCB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPCB_ABI_VERSION 4

typedef enum mpcb_status {
  MPCB_OK = 0,
  MPCB_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
  MPCB_ERR_CUDA = -2,      /* CUDA runtime error (message has the CUDA string) */
  MPCB_ERR_NOMEM = -3,
  MPCB_ERR_STATE = -4      /* call sequence error (e.g. solve before batch_alloc) */
} mpcb_status;

/* Solver status per lane: the values of OSQP's info.status_val (osqp/include/constants.h). */
#define MPCB_QP_SOLVED 1
#define MPCB_QP_SOLVED_INACCURATE 2
#define MPCB_QP_PRIMAL_INFEASIBLE_INACCURATE 3
#define MPCB_QP_MAX_ITER -2
#define MPCB_QP_PRIMAL_INFEASIBLE -3
#define MPCB_QP_UNSOLVED -10

/* Controller codes of SimRun.ctrlr_seq (src/trajectorySimulate.py:378-385). */
#define MPCB_CTRL_MPC 1
#define MPCB_CTRL_FAILSAFE 2
#define MPCB_CTRL_DEADBEAT 3

#define MPCB_EST_UKF 0
#define MPCB_EST_KF 1

typedef struct mpcb_handle mpcb_handle;

/* Constant tables of one problem family, built on the host (mpc_arpo_project_b200/problem.py).
 * All matrices row-major double.  Replaces the setup half of the reference simulators
 * (src/trajectorySimulate.py:73-245) plus OSQP's scale_data / KKT factorisation. */
typedef struct mpcb_problem {
  int32_t Nx, Nc, Nb;            /* horizons, MPCParams (src/mpcsim.py:154-156) */
  int32_t n, m;                  /* n = 4(Nx+1)+7Nc+2 variables, m = 9(Nx+1)+7Nc+2 rows */
  int32_t in_track, delta_v, is_reject, has_noise;   /* SimConditions flags (src/mpcsim.py:59-73) */
  int32_t noise_length;          /* Noise.noise_length, control steps per disturbance draw */
  int32_t estimator;             /* MPCB_EST_UKF: range/bearing UKF of src/trajectorySimulate.py:121-130, 329-337 (default);
                                    MPCB_EST_KF: the linear Kalman filter on position measurements of the reference's
                                    prototype misc/MPCrendezKALMANdisturb.py:261-266, same call timing */
  /* OSQP settings as the reference leaves them (defaults; :245) */
  double rho0, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, adaptive_rho_tolerance;
  int32_t max_iter, check_termination, adaptive_rho, adaptive_rho_interval;
  /* plant, estimator, failsafe (:73-118, :180-187, :272-275) */
  double Ad[16], Bd[8], Ao[36], Bou[12], Qw[36], Kpf[8], Kif[2], xr[4];
  double umax0, r_p, r_tol, suc_dist, suc_ang_deg, mean_mtn, T;
  /* Ruiz-equilibrated QP (OSQP scale_data) */
  const double *P_s;             /* [n*n] */
  const double *q_s;             /* [n]   */
  const double *A_s;             /* [m*n] sign variant 0 (C1 = C2 = +1) */
  const double *l_s, *u_s;       /* [m] scaled bound templates */
  const double *D, *E;           /* [n], [m] */
  double c;
  const int32_t *ctype;          /* [m] -1 free, 0 inequality, 1 equality */
  /* spectral KKT operator per sign variant: M(rho)^-1 = V diag(1/(1+rho*lam)) V' */
  const double *V;               /* [4*n*n] */
  const double *lam;             /* [4*n]   */
  /* Debris-avoidance lanes (src/mpcsim.py:99-123, src/simhelpers.py:48-64,80-134): the half-plane row of every
   * LOS block changes with the estimate each step, so these problems run on the per-lane path, which re-does
   * OSQP's scaling and factorisation on the device every control step and needs the UNSCALED data.  With
   * has_debris != 0 the scaled / spectral tables above may be NULL. */
  int32_t has_debris, scaling;   /* scaling = OSQP `scaling` setting (Ruiz passes, default 10) */
  double debris_center[2], debris_side, debris_detect;
  double debris_verts[8];        /* Debris.constructVertArr() rows, rotated [1,2,3,0] for in-track runs (:52-53) */
  double K_dead[8], Ki_dead[2];  /* deadbeat avoidance law K_total, K_i (src/trajectorySimulate.py:190-203) */
  const double *P_u, *q_u, *A_u, *l_u, *u_u;   /* unscaled P [n*n], q [n], A [m*n] (C1 = C2 = +1, slope = 0), l, u [m] (+-inf allowed) */
} mpcb_problem;

/* Per-trajectory outputs of a simulation; any pointer may be NULL (not recorded).
 * T1 = nsteps+1 for the discrete simulator, n_samples+1 for the continuous one. */
typedef struct mpcb_sim_out {
  int32_t *i_term;       /* [B]  SimRun.i_term (control steps; substeps for the continuous sim) */
  int32_t *is_success;   /* [B]  SimRun.isSuccess (:369-376) */
  double *final_dist;    /* [B]  ||x_true[:, i_term-1] - xr||_2 (test/disturbRejComp.py:87-88) */
  double *x_true;        /* [4][T1][B] */
  double *x_est;         /* [6][T1][B] (x/y swapped for in-track runs, simhelpers.py:72) */
  double *ctrl;          /* [2][T1][B] SimRun.ctrl_hist */
  uint8_t *ctrlr_seq;    /* [T1-1][B]  0 = not reached */
  int8_t *status;        /* [T1-1][B]  OSQP status_val per solve */
  int16_t *iters;        /* [T1-1][B]  ADMM iterations per solve */
  double *u_raw;         /* [2][T1-1][B] selected control before the norm clip (:314) */
  int32_t *ukf_clamped;  /* [B]  1 if a UKF Cholesky pivot was <= 0 and clamped (the reference raises LinAlgError there) */
  /* continuous simulator only: telemetry at EVERY substep, the reference's own array shapes
   * (src/trajectorySimulateC.py:284-292, 414-443); NS = n_sub_total.  Meant for small batches. */
  double *x_true_sub;    /* [4][NS][B]  xtrueP */
  double *ctrl_sub;      /* [2][NS][B]  ctrls */
  uint8_t *ctrlr_sub;    /* [NS][B]     controllerSeq codes 0/1/2 */
  /* solver telemetry beyond the reference's SimRun (ABI 4) */
  double *rho;           /* [T1-1][B]  the lane's ADMM step size after solve i (OSQP's adaptive rho; NaN = not reached) */
} mpcb_sim_out;

/* Run counters filled by the simulate / qp_solve calls (for bench.py's gpu_launches etc.). */
typedef struct mpcb_counters {
  int64_t qp_solves;        /* closed-loop QP solves = live trajectory control steps */
  int64_t admm_iterations;  /* total ADMM iterations over all lanes */
  int64_t kernel_launches;  /* kernels launched by this library */
  int64_t admm_launches;    /* of which: ADMM block kernel */
  int64_t rounds;           /* lockstep rounds (one 25-iteration block per live lane) */
  int64_t flip_lanes;       /* lanes in which OSQP re-types a velocity-bound row as an equality (scaled u - l < 1e-4): modelled by the
                               team / wave+team / per-lane kernels, only counted by the block (Nx = 40) and tile kernels */
  int64_t operator_rebuilds;/* per-trajectory KKT operator rebuilds (rho adapted or velocity signs flipped) */
  double admm_ms;           /* CUDA-event time spent in the ADMM block kernel (if timing enabled) */
  double total_ms;          /* CUDA-event time of the whole call, on the handle's stream */
} mpcb_counters;

int mpcb_abi_version(void);
const char *mpcb_last_error(void);

/* osqp.OSQP().setup(P,q,A,l,u,...) for the whole batch (:242-245): uploads the shared tables. */
int mpcb_create(const mpcb_problem *problem, int device, mpcb_handle **out);
int mpcb_destroy(mpcb_handle *h);

/* Allocate per-lane state for B lanes and cold-start every lane's solver (x = z = y = 0, rho = rho0). */
int mpcb_batch_alloc(mpcb_handle *h, int64_t B);
int mpcb_set_timing(mpcb_handle *h, int enable_kernel_timing);
int mpcb_get_counters(mpcb_handle *h, mpcb_counters *out);
/* Which solver blocks this handle can run (bit mask): 1 = warp-per-lane block kernel, 2 = team kernel (whole closed loop per
 * CTA), 4 = DMMA tile kernel, 8 = multi-RHS wave kernel (DMMA, unscaled sparse phases), 16 = per-lane debris kernel.
 * MPCB_SOLVER = block | team | tile | wave forces one; tests use the mask to know that the forced block really ran. */
int mpcb_solver_blocks(mpcb_handle *h);
/* The CUDA stream (cudaStream_t) the handle launches on, for callers that time with events. */
void *mpcb_stream(mpcb_handle *h);
/* Order the handle's stream after everything enqueued so far on `producer_stream` (a cudaStream_t; NULL = the legacy
 * default stream): cudaEventRecord + cudaStreamWaitEvent, no host synchronisation.  Call it before passing device
 * pointers whose contents are still being produced on another stream (see the stream contract above). */
int mpcb_wait_stream(mpcb_handle *h, void *producer_stream);

/* prob.update(l,u); prob.update(Ax,l,u); res = prob.solve()  (:340-348, :296) for every lane:
 * xhat[6][B] is the estimate the bounds / signs are rebuilt from (simhelpers.py:66-67,124,137);
 * the warm start is whatever the previous call left in the lane.  Outputs: u0[2][B] =
 * res.x[(Nx+1)*nx : +nu], status[B] = res.info.status_val, iters[B] = res.info.iter. */
int mpcb_qp_solve(mpcb_handle *h, int64_t B, const double *xhat, double *u0, int32_t *status, int32_t *iters,
                  int io_on_device);
/* Test seam: read back a lane's scaled ADMM iterates (x[n], z[m], y[m], rho) to the host. */
int mpcb_qp_get_state(mpcb_handle *h, int64_t lane, double *x, double *z, double *y, double *rho);

/* kf.predict(u); kf.update(z)  (:333-335) on x[6][B], P[36][B] in place. */
int mpcb_ukf_step(mpcb_handle *h, int64_t B, double *x, double *P, const double *u, const double *z, int io_on_device);
/* x <- Ad x + Bd u + [w;0]  (:324) on x[4][B]. */
int mpcb_plant_lin_step(mpcb_handle *h, int64_t B, double *x, const double *u, const double *w, int io_on_device);
/* nsub fixed-step RK4 substeps of the nonlinear relative-motion ODE (trajectorySimulateC.py:64-79,
 * 372-380) on x[4][B]: x <- RK4(x, u, dt) + [w;0] per substep. */
int mpcb_plant_rk4(mpcb_handle *h, int64_t B, double *x, const double *u, const double *w, int nsub, double dt,
                   int io_on_device);

/* trajectorySimulate (src/trajectorySimulate.py:17-388) for B lanes in lockstep rounds.
 * x0[4][B]; noise[n_refresh][2][B] are the sigma-scaled position disturbances drawn at step 0 and
 * after every noise_length steps (:268, :351-356); NULL when has_noise == 0. */
int mpcb_simulate_discrete(mpcb_handle *h, int64_t B, int32_t nsteps, const double *x0, const double *noise,
                           int32_t n_refresh, const mpcb_sim_out *out, int io_on_device);

/* trajectorySimulateC (src/trajectorySimulateC.py:17-446): nonlinear plant, RK4 with step T_cont,
 * `ratio` = int(T/T_cont) substeps per control interval, `n_sub_total` = int(T_final/T_cont).
 * noise[n_refresh][2][B]: per-substep additive position disturbance held for `noise_hold_sub`
 * substeps (:296-307).  Telemetry is decimated to the sample instants. */
int mpcb_simulate_continuous(mpcb_handle *h, int64_t B, int32_t n_sub_total, int32_t ratio, double T_cont,
                             const double *x0, const double *noise, int32_t n_refresh, int32_t noise_hold_sub,
                             const mpcb_sim_out *out, int io_on_device);

/* Final statistics of the last simulation, reduced over the batch on the device
 * (test/disturbRejComp.py:89-100, test/saved_runs/success_rates_test.py:66-75):
 * stats[0]=sum final_dist, [1]=sum final_dist^2, [2]=#success, [3]=#lanes, [4]=sum i_term,
 * [5]=qp solves, [6]=admm iterations, [7]=flip lanes, [8]=lanes with a clamped UKF Cholesky pivot,
 * [9]=lanes that terminated before the last step.  These MPCB_NSTATS doubles are what ranks
 * all-reduce (sum) over NCCL. */
#define MPCB_NSTATS 10
int mpcb_stats(mpcb_handle *h, int64_t B, double *stats, int io_on_device);
/* The one collective of a multi-GPU Monte-Carlo run: sum the MPCB_NSTATS statistics of the last simulation
 * over the ranks of `nccl_comm` (an ncclComm_t the caller created with one rank per GPU), on the handle's stream, and
 * return the global vector in stats (host pointer).  libnccl.so.2 is resolved at run time (dlopen): single-GPU users
 * need no NCCL.  The reductions are sums of per-rank values that are themselves fixed-order sums, so the result does not
 * depend on timing. */
int mpcb_allreduce_stats(mpcb_handle *h, void *nccl_comm, double *stats);

/* Disturbance draws on the device: noise[n_refresh][2][B] = (sigma_x, sigma_y) .* N(0,1) from Philox4x32-10, one block per
 * (lane + lane_offset, refresh), key = seed -- the reference's model `sigMat @ random.normal(0,1,4)` per noise_length steps
 * (src/trajectorySimulate.py:268, 351-356: four draws, the two position entries used), not numpy's stream.  raw (optional,
 * may be NULL): the generator's words [n_refresh][4][B] for known-answer tests.  lane_offset = rank * B keeps the ranks of a
 * multi-GPU run on disjoint streams. */
int mpcb_noise_fill(mpcb_handle *h, int64_t B, int32_t n_refresh, double sigma_x, double sigma_y, uint64_t seed,
                    uint64_t lane_offset, double *noise, uint32_t *raw, int io_on_device);

/* Bench utility (no reference counterpart): float64 peak of the device in TFLOP/s, measured with a
 * register-resident DFMA loop (use_dmma = 0) or mma.sync.m8n8k4.f64 loop (use_dmma = 1); the roofline
 * denominators for this float64 path (MEASURED_PEAKS.json has none). */
int mpcb_measure_fp64_peak(int device, int use_dmma, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* MPCB_H */
