// libmpcb.so -- C ABI (include/mpcb.h) over the sm_100a kernels in admm.cuh / sim.cuh.
// Host side only: table upload, per-batch state, the lockstep round loop, I/O staging.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/mpcb.h"
#include "admm.cuh"
#include "common.cuh"
#include "sim.cuh"
#include "team.cuh"
#include "generic.cuh"
#include "tile.cuh"
#include "wave.cuh"
#include <stdlib.h>
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>

// NVTX ranges around the entry points: visible in Nsight Systems timelines, free otherwise
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      char b_[512];                                                                                      \
      snprintf(b_, sizeof b_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));      \
      return fail(e_ == cudaErrorMemoryAllocation ? MPCB_ERR_NOMEM : MPCB_ERR_CUDA, b_);                 \
    }                                                                                                    \
  } while (0)

#define RC(call)                 \
  do {                           \
    int rc_ = (call);            \
    if (rc_ != MPCB_OK) return rc_; \
  } while (0)

typedef void (*admm_fn)(const AdmmArgs);

struct KernelChoice {
  admm_fn fn = nullptr;
  int max_warps = 0;
};

template <int NXS, int MZS, bool VS, int WARPS>
static KernelChoice pick() {
  KernelChoice k;
  k.fn = admm_block_kernel<NXS, MZS, VS, WARPS>;
  k.max_warps = WARPS;
  return k;
}

// Instantiations cover the reference's Nc = Nb = 5 problem family at the horizons the reference
// scripts and BASELINE.json use: Nx = 10, 20, 30, 40 (n = 4Nx+41, m = 9Nx+46).
static KernelChoice choose_kernel(int nxs, int mzs, bool vs) {
#define TRY(NX_, MZ_, W_)                                                      \
  if (nxs == NX_ && mzs == MZ_) return vs ? pick<NX_, MZ_, true, W_>() : pick<NX_, MZ_, false, W_>();
  TRY(3, 5, 16) TRY(4, 8, 8) TRY(6, 10, 8) TRY(7, 13, 8)
#undef TRY
  return KernelChoice();
}

struct HostProblem {
  mpcb_problem p;
  std::vector<double> P_s, q_s, A_s, l_s, u_s, D, E, V, lam;
  std::vector<int32_t> ctype;
  std::vector<double> P_u, q_u, A_u, l_u, u_u;   // unscaled data (per-lane path)
};

struct StageBuf {
  void *ptr = nullptr;
  size_t cap = 0;
};

struct mpcb_handle {
  std::vector<StageBuf> stage_pool;   // cached device staging buffers for host-pointer calls
  int device = 0;
  cudaStream_t stream = nullptr;
  HostProblem hp;
  SimConst sc;
  // constant tables
  BlobHdr hdr;
  unsigned char *d_blob[4] = {nullptr, nullptr, nullptr, nullptr};
  double *d_V[4] = {nullptr, nullptr, nullptr, nullptr};
  bool vsmem = false;
  KernelChoice kern;
  int warps = 8;
  size_t smem_bytes = 0;
  double qn_unscaled = 0, qn_scaled = 0;
  int num_sms = 148;
  // team kernel (n = 81 family): whole closed loop per CTA, operator in registers
  bool team_ok = false;
  TeamHdr thdr;
  unsigned char *d_tblob = nullptr;
  double *d_Vk[4] = {nullptr, nullptr, nullptr, nullptr};
  double *d_lam = nullptr;
  int *d_queue = nullptr;
  // per-lane operator cache of the tensor-memory team kernel (list mode), allocated on first use
  double *scache = nullptr, *scache_rho = nullptr;
  double *ocache = nullptr;            // per-team operator cache of the whole-loop modes (TeamArgs::ocache)
  int *scache_var = nullptr;
  bool scache_tried = false, team_tm = false;
  int team_threads = 256;
  int visit_iters = -1;                // list mode: iterations per lane visit; -1 = sized per round, 0 = unlimited (MPCB_VISIT_ITERS)
  size_t team_smem = 0;
  const void *team_fn_ptr = nullptr;
  int team_ctas = 0;
  // tile kernel (DMMA, 8 lanes per warp): the round-based solver block for large batches of the n = 81 family
  bool tile_ok = false;
  TileHdr tile_hdr;
  unsigned char *d_tile_blob[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t tile_smem = 0;
  int tile_warps = 4;
  int64_t tile_min_lanes = INT64_MAX;  // the tile kernel is opt-in (MPCB_SOLVER=tile): the team kernel is faster at every batch size measured so far
  // wave kernel (multi-RHS DMMA, 8 lanes per warp, unscaled sparse phases): round-based solver block of the n = 81 family
  bool wave_ok = false;
  WaveHdr wave_hdr;
  WaveConst wave_k;
  unsigned char *d_wave_blob[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t wave_smem = 0, wave_smem4 = 0;
  int wave_warps_now = WAVE_WARPS;  // CTA size of the current simulate call's wave rounds (8 or 4 warps)
  int64_t wave_min_lanes = 16384;      // default solver block of the round-based simulators from this batch size on
  // per-lane path (debris lanes)
  bool generic_ok = false;
  GenArgs gproto;
  std::vector<void *> gen_bufs;
  int gen_grid = 0;
  size_t gen_smem = 0;
  // batch state
  int64_t B = 0;
  double *xs = nullptr, *zs = nullptr, *ys = nullptr, *rho = nullptr, *par = nullptr, *u0 = nullptr;
  int *iter = nullptr, *status = nullptr, *flip = nullptr;
  uint8_t *lane_state = nullptr;
  int *cnt = nullptr;    // [2][4]
  int *list = nullptr;   // [2][4][B]
  LaneSim ls;
  double *lane_f64 = nullptr;   // backing store of LaneSim doubles
  double *lane_fd = nullptr;    // [B] final distance of every lane (fixed-order statistics)
  int *lane_i32 = nullptr;      // backing store of LaneSim ints
  unsigned long long *d_tot = nullptr;   // [0] admm iterations, [1] qp solves, [2] operator rebuilds, [3] spare
  double *d_stats = nullptr;             // [MPCB_NSTATS]
  int *h_cnt = nullptr;                  // pinned [4]
  bool sim_done = false;
  int64_t last_sim_iterations = 0;
  // timing
  bool timing = false;
  std::vector<cudaEvent_t> ev_pool;
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_wait = nullptr;
  mpcb_counters ctr;
};

// ------------------------------------------------------------------------------------------
// grouped ELL: rows in groups of 32 (row = lane + 32*s); a group stores width*32 entries [e][lane].
struct Ell {
  std::vector<double> vals;
  std::vector<uint16_t> cols;
  std::vector<int2> grp;
};

static Ell build_ell(const double *M, int rows, int cols, bool transpose, int ngroups) {
  // logical matrix L(r, c) = transpose ? M[c*rows_in + r] : M[r*cols + c]; `rows`/`cols` are L's dims
  Ell e;
  auto at = [&](int r, int c) { return transpose ? M[(size_t)c * rows + r] : M[(size_t)r * cols + c]; };
  int off = 0;
  for (int g = 0; g < ngroups; ++g) {
    int width = 0;
    for (int ln = 0; ln < 32; ++ln) {
      const int r = g * 32 + ln;
      if (r >= rows) continue;
      int nz = 0;
      for (int c = 0; c < cols; ++c) nz += (at(r, c) != 0.0);
      width = std::max(width, nz);
    }
    e.grp.push_back(make_int2(off, width));
    e.vals.resize(off + width * 32, 0.0);
    e.cols.resize(off + width * 32, 0);
    for (int ln = 0; ln < 32; ++ln) {
      const int r = g * 32 + ln;
      if (r >= rows) continue;
      int k = 0;
      for (int c = 0; c < cols; ++c) {
        const double v = at(r, c);
        if (v != 0.0) {
          e.vals[off + k * 32 + ln] = v;
          e.cols[off + k * 32 + ln] = (uint16_t)c;
          ++k;
        }
      }
    }
    off += width * 32;
  }
  return e;
}

static int align16(int x) { return (x + 15) & ~15; }

static void variant_matrix(const HostProblem &hp, int v, std::vector<double> &A) {
  const mpcb_problem &p = hp.p;
  A = hp.A_s;
  const int nX = 4 * (p.Nx + 1);
  for (int k = 0; k <= p.Nx; ++k) {
    const int r = nX + 5 * k + 3;
    if (v & 1) A[(size_t)r * p.n + 4 * k + 2] *= -1.0;
    if (v & 2) A[(size_t)r * p.n + 4 * k + 3] *= -1.0;
  }
}

static int build_tables(mpcb_handle *h) {
  const HostProblem &hp = h->hp;
  const mpcb_problem &p = hp.p;
  const int n = p.n, m = p.m;
  const int nxs = (n + 31) / 32, mzs = (m + 31) / 32;
  std::vector<std::vector<unsigned char>> blobs(4);
  BlobHdr hdr;
  memset(&hdr, 0, sizeof hdr);
  // Does V fit in shared memory next to the rest?  Decide after the first variant's layout is known.
  for (int pass = 0; pass < 2; ++pass) {
    for (int v = 0; v < 4; ++v) {
      std::vector<double> A;
      variant_matrix(hp, v, A);
      Ell eA = build_ell(A.data(), m, n, false, mzs);
      Ell eAT = build_ell(A.data(), n, m, true, nxs);
      Ell eP = build_ell(hp.P_s.data(), n, n, false, nxs);
      int off = 0;
      auto take = [&](int bytes) {
        const int o = off;
        off = align16(off + bytes);
        return o;
      };
      BlobHdr hv;
      memset(&hv, 0, sizeof hv);
      hv.off_lam = take(8 * n);
      hv.off_q = take(8 * n);
      hv.off_D = take(8 * n);
      hv.off_Dinv = take(8 * n);
      hv.off_E = take(8 * m);
      hv.off_Einv = take(8 * m);
      hv.off_lt = take(8 * m);
      hv.off_ut = take(8 * m);
      hv.off_Av = take(8 * (int)eA.vals.size());
      hv.off_ATv = take(8 * (int)eAT.vals.size());
      hv.off_Pv = take(8 * (int)eP.vals.size());
      hv.off_Ac = take(2 * (int)eA.cols.size());
      hv.off_ATc = take(2 * (int)eAT.cols.size());
      hv.off_Pc = take(2 * (int)eP.cols.size());
      hv.off_Ag = take(8 * (int)eA.grp.size());
      hv.off_ATg = take(8 * (int)eAT.grp.size());
      hv.off_Pg = take(8 * (int)eP.grp.size());
      hv.off_flags = take(m);
      hv.off_V = (pass == 1 && h->vsmem) ? take(8 * n * n) : -1;
      hv.total = off;
      if (v == 0) hdr = hv;
      else if (memcmp(&hdr, &hv, sizeof hv) != 0)
        return fail(MPCB_ERR_INVALID, "sign variants do not share one sparsity layout");
      if (pass == 0) continue;
      std::vector<unsigned char> &b = blobs[v];
      b.assign(hv.total, 0);
      auto putd = [&](int o, const double *src, size_t cnt) { memcpy(b.data() + o, src, cnt * 8); };
      std::vector<double> Dinv(n), Einv(m);
      for (int j = 0; j < n; ++j) Dinv[j] = 1.0 / hp.D[j];
      for (int i = 0; i < m; ++i) Einv[i] = 1.0 / hp.E[i];
      putd(hv.off_lam, hp.lam.data() + (size_t)v * n, n);
      putd(hv.off_q, hp.q_s.data(), n);
      putd(hv.off_D, hp.D.data(), n);
      putd(hv.off_Dinv, Dinv.data(), n);
      putd(hv.off_E, hp.E.data(), m);
      putd(hv.off_Einv, Einv.data(), m);
      putd(hv.off_lt, hp.l_s.data(), m);
      putd(hv.off_ut, hp.u_s.data(), m);
      putd(hv.off_Av, eA.vals.data(), eA.vals.size());
      putd(hv.off_ATv, eAT.vals.data(), eAT.vals.size());
      putd(hv.off_Pv, eP.vals.data(), eP.vals.size());
      memcpy(b.data() + hv.off_Ac, eA.cols.data(), eA.cols.size() * 2);
      memcpy(b.data() + hv.off_ATc, eAT.cols.data(), eAT.cols.size() * 2);
      memcpy(b.data() + hv.off_Pc, eP.cols.data(), eP.cols.size() * 2);
      memcpy(b.data() + hv.off_Ag, eA.grp.data(), eA.grp.size() * 8);
      memcpy(b.data() + hv.off_ATg, eAT.grp.data(), eAT.grp.size() * 8);
      memcpy(b.data() + hv.off_Pg, eP.grp.data(), eP.grp.size() * 8);
      for (int i = 0; i < m; ++i) {
        uint8_t f = 0;
        if (hp.l_s[i] < -1e30 * 1e-4) f |= 1;
        if (hp.u_s[i] > 1e30 * 1e-4) f |= 2;
        if (hp.ctype[i] == 1) f |= 4;
        if (hp.ctype[i] == -1) f |= 8;
        b[hv.off_flags + i] = f;
      }
      if (hv.off_V >= 0) putd(hv.off_V, hp.V.data() + (size_t)v * n * n, (size_t)n * n);
    }
    if (pass == 0) {
      // per-warp scratch (m + 2n doubles) + barrier; aim for >= 8 warps with V resident
      const int fixed = hdr.total + 16;
      const int withV = fixed + 8 * n * n;
      h->vsmem = (withV + 8 * 8 * (m + 2 * n)) <= 220 * 1024;
    }
  }
  h->hdr = hdr;
  h->kern = choose_kernel(nxs, mzs, h->vsmem);
  if (!h->kern.fn) {
    char b[160];
    snprintf(b, sizeof b, "no ADMM kernel instantiation for n=%d m=%d (Nc = 5 with Nx in {10, 20, 30, 40} supported)", n, m);
    return fail(MPCB_ERR_INVALID, b);
  }
  // warps per CTA: as many as fit in shared memory, capped by the instantiation's launch bound
  const int per_warp = 8 * (m + 2 * n);
  int w = (int)((220 * 1024 - (hdr.total + 16)) / per_warp);
  w = std::min(w, h->kern.max_warps);
  if (w < 1) return fail(MPCB_ERR_INVALID, "problem too large for shared memory staging");
  h->warps = w;
  for (int v = 0; v < 4; ++v) {
    CK(cudaMalloc(&h->d_blob[v], hdr.total));
    CK(cudaMemcpy(h->d_blob[v], blobs[v].data(), hdr.total, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&h->d_V[v], (size_t)8 * n * n));
    CK(cudaMemcpy(h->d_V[v], hp.V.data() + (size_t)v * n * n, (size_t)8 * n * n, cudaMemcpyHostToDevice));
  }
  double qu = 0, qs = 0;
  for (int j = 0; j < n; ++j) {
    qu = std::max(qu, fabs(hp.q_s[j] / hp.D[j]));
    qs = std::max(qs, fabs(hp.q_s[j]));
  }
  h->qn_unscaled = qu;
  h->qn_scaled = qs;
  return MPCB_OK;
}

static void set_warps(mpcb_handle *h, int w) {
  h->warps = w;
  h->smem_bytes = (size_t)h->hdr.total + 16 + (size_t)w * 8 * (h->hp.p.m + 2 * h->hp.p.n);
}


// ------------------------------------------------------------------------------------------
// Team kernel tables: ELL in "team layout" (row r -> thread r % TEAM, slot r / TEAM; entry e of a
// slot at [(base + e) * TEAM + thread]), the sign-patch list, k-major V per variant.
typedef void (*team_fn)(const TeamArgs);
struct TeamShape { int n, m, wa, wat2, watx, wp, ss, threads, ctas; team_fn fn; };
static const TeamShape kTeamShapes[] = {
  // Nx = 10 (BASELINE configs 2, 3): operator in tensor memory (84 columns per thread, 2 teams per SM), A / A' rows in registers
  {81, 136, 8, 6, 0, 4, 0, 256, 2, team_kernel<81, 136, 8, 6, 0, 4, 0, true, 256, 2>},
  // same family, operator in registers + 8 entries per thread in shared memory (MPCB_TEAM=regs)
  {81, 136, 8, 6, 0, 4, 8, 256, 2, team_kernel<81, 136, 8, 6, 0, 4, 8, false, 256, 2>},
  // Nx = 20 (config 4): 124 columns per thread, two warps per lane quarter -> 248 of a 256-column allocation, 2 teams per SM
  {121, 226, 8, 6, 5, 4, 0, 256, 2, team_kernel<121, 226, 8, 6, 5, 4, 0, true, 256, 2>},
  // Nx = 30 (config 5): 164 columns per thread, three warps per lane quarter -> 492 of 512 columns, one 384-thread team per SM
  {161, 316, 8, 6, 10, 4, 0, 384, 1, team_kernel<161, 316, 8, 6, 10, 4, 0, true, 384, 1>},
};
static const TeamShape *team_shape_for(int n, int m) {
  const char *e = getenv("MPCB_TEAM");
  const bool regs = e && strcmp(e, "regs") == 0;
  for (const TeamShape &t : kTeamShapes)
    if (t.n == n && t.m == m && (t.ss != 0) == regs) return &t;
  return nullptr;
}

static int build_team_tables(mpcb_handle *h) {
  const HostProblem &hp = h->hp;
  const mpcb_problem &p = hp.p;
  const int n = p.n, m = p.m;
  const TeamShape *ts = team_shape_for(n, m);
  if (!ts || getenv("MPCB_FORCE_V1")) return MPCB_OK;
  const int mp = (m + 1) & ~1, nct = ((2 * n + 31) / 32) * 32, npair = nct / 2;
  // nnz lists: rows of A, columns of A (= rows of A'), rows of P
  std::vector<std::vector<int>> rowsA(m), colsA(n), rowsP(n);
  for (int r = 0; r < m; ++r)
    for (int c = 0; c < n; ++c)
      if (hp.A_s[(size_t)r * n + c] != 0.0) {
        rowsA[r].push_back(c);
        colsA[c].push_back(r);
      }
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c)
      if (hp.P_s[(size_t)r * n + c] != 0.0) rowsP[r].push_back(c);
  // hand out rows / columns in order of decreasing nnz so every warp gets its own ELL width
  std::vector<int> rowmap(m), colmap(n), rowpos(m), colpos(n);
  for (int r = 0; r < m; ++r) rowmap[r] = r;
  for (int c = 0; c < n; ++c) colmap[c] = c;
  std::stable_sort(rowmap.begin(), rowmap.end(), [&](int x, int y) { return rowsA[x].size() > rowsA[y].size(); });
  std::stable_sort(colmap.begin(), colmap.end(), [&](int x, int y) { return colsA[x].size() > colsA[y].size(); });
  for (int t = 0; t < m; ++t) rowpos[rowmap[t]] = t;
  for (int q = 0; q < n; ++q) colpos[colmap[q]] = q;
  TeamHdr hd;
  memset(&hd, 0, sizeof hd);
  for (int t = 0; t < m; ++t) {
    const int w = (int)rowsA[rowmap[t]].size();
    if (w > ts->wa || w > 8) return MPCB_OK;                 // wider than the instantiation: stay on the generic path
    hd.wA[t / 32] = std::max(hd.wA[t / 32], w);
  }
  for (int q = 0; q < n; ++q) {
    const int w2 = ((int)colsA[colmap[q]].size() + 1) / 2;
    if (w2 > ts->wat2 + ts->watx || w2 > 16) return MPCB_OK;
    hd.wAT2[(2 * q) / 32] = std::max(hd.wAT2[(2 * q) / 32], w2);
    if ((int)rowsP[colmap[q]].size() > ts->wp) return MPCB_OK;
  }
  const int wat_all = ts->wat2 + ts->watx;
  std::vector<double> Av((size_t)ts->wa * mp, 0.0), ATv((size_t)wat_all * nct, 0.0), Pv((size_t)ts->wp * npair, 0.0);
  std::vector<uint16_t> Ac((size_t)8 * mp, 0), ATc((size_t)8 * nct, 0), ATc2((size_t)8 * nct, 0), Pc((size_t)ts->wp * npair, 0), rmap(mp, 0),
      cmap(npair, 0);
  for (int t = 0; t < m; ++t) {
    const int r = rowmap[t];
    rmap[t] = (uint16_t)r;
    for (size_t e = 0; e < rowsA[r].size(); ++e) {
      Av[e * mp + t] = hp.A_s[(size_t)r * n + rowsA[r][e]];
      Ac[(size_t)t * 8 + e] = (uint16_t)rowsA[r][e];
    }
  }
  for (int q = 0; q < n; ++q) {
    const int c = colmap[q];
    cmap[q] = (uint16_t)c;
    for (size_t k = 0; k < colsA[c].size(); ++k) {
      const int tid = 2 * q + (int)(k & 1);
      ATv[(k / 2) * nct + tid] = hp.A_s[(size_t)colsA[c][k] * n + c];
      if (k / 2 < 8) ATc[(size_t)tid * 8 + k / 2] = (uint16_t)colsA[c][k];
      else ATc2[(size_t)tid * 8 + k / 2 - 8] = (uint16_t)colsA[c][k];
    }
    for (size_t e = 0; e < rowsP[c].size(); ++e) {
      Pv[e * npair + q] = hp.P_s[(size_t)c * n + rowsP[c][e]];
      Pc[e * npair + q] = (uint16_t)rowsP[c][e];
    }
  }
  std::vector<int4> patch;
  const int nX = 4 * (p.Nx + 1);
  for (int k = 0; k <= p.Nx; ++k) {
    const int r = nX + 5 * k + 3;
    for (int which = 1; which <= 2; ++which) {
      const int c = 4 * k + 1 + which;
      int ia = -1, iat = -1;
      for (size_t e = 0; e < rowsA[r].size(); ++e)
        if (rowsA[r][e] == c) ia = (int)(e * mp + rowpos[r]);
      for (size_t kk = 0; kk < colsA[c].size(); ++kk)
        if (colsA[c][kk] == r) iat = (int)((kk / 2) * nct + 2 * colpos[c] + (kk & 1));
      if (ia < 0 || iat < 0) return fail(MPCB_ERR_INVALID, "velocity-sign entry missing from A");
      patch.push_back(make_int4(ia, iat, which, 0));
    }
  }
  int off = 0;
  auto take = [&](size_t bytes) {
    const int o = off;
    off = align16(off + (int)bytes);
    return o;
  };
  hd.off_q = take(8 * n); hd.off_D = take(8 * n); hd.off_Dinv = take(8 * n);
  hd.off_E = take(8 * m); hd.off_Einv = take(8 * m); hd.off_lt = take(8 * m); hd.off_ut = take(8 * m);
  hd.off_Av = take(8 * Av.size()); hd.off_ATv = take(8 * ATv.size()); hd.off_Pv = take(8 * Pv.size());
  hd.off_Ac = take(2 * Ac.size()); hd.off_ATc = take(2 * ATc.size()); hd.off_ATc2 = take(2 * ATc2.size()); hd.off_Pc = take(2 * Pc.size());
  hd.off_rowmap = take(2 * rmap.size()); hd.off_colmap = take(2 * cmap.size());
  hd.off_flags = take(m);
  {   // shape of the rows OSQP may re-type (team.cuh tm_retype_operator): +-1 on the stage's two velocities, at most one more entry
    hd.r3ok = (p.Nb + 1 <= 8) ? 1 : 0;
    for (int k = 0; k < 8; ++k) { hd.r3c[k] = -1; hd.r3v[k] = 0.0; }
    for (int k = 0; k <= p.Nb && hd.r3ok; ++k) {
      const int r = 4 * (p.Nx + 1) + 5 * k + 3;
      int extra = 0;
      for (int j = 0; j < n; ++j) {
        const double v = hp.A_s[(size_t)r * n + j];
        if (v == 0.0 || j == 4 * k + 2 || j == 4 * k + 3) continue;
        hd.r3c[k] = j;
        hd.r3v[k] = v;
        ++extra;
      }
      if (extra > 1 || hp.A_s[(size_t)r * n + 4 * k + 2] == 0.0 || hp.A_s[(size_t)r * n + 4 * k + 3] == 0.0) hd.r3ok = 0;
    }
  }
  hd.off_patch = take(16 * patch.size());
  hd.n_patch = (int)patch.size();
  hd.total = off;
  std::vector<unsigned char> b(hd.total, 0);
  auto putd = [&](int o, const double *src, size_t cnt) { memcpy(b.data() + o, src, cnt * 8); };
  std::vector<double> Dinv(n), Einv(m);
  for (int j = 0; j < n; ++j) Dinv[j] = 1.0 / hp.D[j];
  for (int i = 0; i < m; ++i) Einv[i] = 1.0 / hp.E[i];
  putd(hd.off_q, hp.q_s.data(), n); putd(hd.off_D, hp.D.data(), n); putd(hd.off_Dinv, Dinv.data(), n);
  putd(hd.off_E, hp.E.data(), m); putd(hd.off_Einv, Einv.data(), m);
  putd(hd.off_lt, hp.l_s.data(), m); putd(hd.off_ut, hp.u_s.data(), m);
  putd(hd.off_Av, Av.data(), Av.size());
  putd(hd.off_ATv, ATv.data(), ATv.size());
  putd(hd.off_Pv, Pv.data(), Pv.size());
  memcpy(b.data() + hd.off_Ac, Ac.data(), Ac.size() * 2);
  memcpy(b.data() + hd.off_ATc, ATc.data(), ATc.size() * 2);
  memcpy(b.data() + hd.off_ATc2, ATc2.data(), ATc2.size() * 2);
  memcpy(b.data() + hd.off_Pc, Pc.data(), Pc.size() * 2);
  memcpy(b.data() + hd.off_rowmap, rmap.data(), rmap.size() * 2);
  memcpy(b.data() + hd.off_colmap, cmap.data(), cmap.size() * 2);
  for (int i = 0; i < m; ++i) {
    uint8_t f = 0;
    if (hp.l_s[i] < -1e30 * 1e-4) f |= 1;
    if (hp.u_s[i] > 1e30 * 1e-4) f |= 2;
    if (hp.ctype[i] == 1) f |= 4;
    if (hp.ctype[i] == -1) f |= 8;
    b[hd.off_flags + i] = f;
  }
  memcpy(b.data() + hd.off_patch, patch.data(), patch.size() * 16);
  CK(cudaMalloc(&h->d_tblob, hd.total));
  CK(cudaMemcpy(h->d_tblob, b.data(), hd.total, cudaMemcpyHostToDevice));
  const int half = ((n + 3) / 4) * 2, np2 = 2 * half;
  for (int v = 0; v < 4; ++v) {
    std::vector<double> vk((size_t)n * np2, 0.0);
    const double *V = hp.V.data() + (size_t)v * n * n;
    for (int k = 0; k < n; ++k)
      for (int j = 0; j < n; ++j) vk[(size_t)k * np2 + j] = V[(size_t)j * n + k];
    CK(cudaMalloc(&h->d_Vk[v], vk.size() * 8));
    CK(cudaMemcpy(h->d_Vk[v], vk.data(), vk.size() * 8, cudaMemcpyHostToDevice));
  }
  CK(cudaMalloc(&h->d_lam, (size_t)4 * n * 8));
  CK(cudaMemcpy(h->d_lam, hp.lam.data(), (size_t)4 * n * 8, cudaMemcpyHostToDevice));
  h->thdr = hd;
  const int nw = ts->threads / 32;
  h->team_smem = (size_t)hd.total + 8 * (size_t)(5 * mp + 3 * np2 + 16 * nw + ts->ss * nct) + ((sizeof(UkfScratch) + 15) & ~15) +
                 sizeof(LaneCtx) + 16;
  h->team_ctas = ts->ctas;
  h->team_threads = ts->threads;
  h->team_tm = ts->ss == 0;
  if (h->team_tm && n > 100 && !getenv("MPCB_NO_OCACHE")) {   // (team.cuh: OC)       // [grid][TEAM_OPC] operators; stays L2-resident for the small families
    const size_t per_op = (size_t)((n + 3) / 4) * nct * 16;
    if (cudaMalloc(&h->ocache, per_op * TEAM_OPC * (size_t)h->num_sms * ts->ctas) != cudaSuccess) {
      cudaGetLastError();
      h->ocache = nullptr;
    }
  }
  h->team_fn_ptr = (const void *)ts->fn;
  CK(cudaFuncSetAttribute(h->team_fn_ptr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->team_smem));
  h->team_ok = true;
  return MPCB_OK;
}

static void fill_team_args(mpcb_handle *h, TeamArgs &ta, int mode) {
  const mpcb_problem &p = h->hp.p;
  memset(&ta, 0, sizeof ta);
  ta.hdr = h->thdr;
  ta.blob = h->d_tblob;
  for (int v = 0; v < 4; ++v) ta.Vk[v] = h->d_Vk[v];
  ta.lam = h->d_lam;
  ta.n = p.n; ta.m = p.m; ta.nX = 4 * (p.Nx + 1); ta.Nb = p.Nb; ta.uoff = 4 * (p.Nx + 1); ta.B = (int)h->B;
  ta.sigma = p.sigma; ta.alpha = p.alpha; ta.eps_abs = p.eps_abs; ta.eps_rel = p.eps_rel; ta.eps_pinf = p.eps_prim_inf;
  ta.adapt_tol = p.adaptive_rho_tolerance; ta.cinv = 1.0 / p.c; ta.qn_unscaled = h->qn_unscaled; ta.qn_scaled = h->qn_scaled;
  ta.rho0 = std::min(std::max(p.rho0, MPCB_RHO_MIN), MPCB_RHO_MAX);
  ta.check_every = p.check_termination; ta.adaptive = p.adaptive_rho; ta.adapt_interval = std::max(1, p.adaptive_rho_interval);
  ta.max_iter = p.max_iter;
  ta.mode = mode;
  ta.sc = h->sc;
  ta.xs = h->xs; ta.zs = h->zs; ta.ys = h->ys; ta.rho = h->rho; ta.u0 = h->u0;
  ta.iter = h->iter; ta.status = h->status; ta.flip = h->flip;
  ta.queue = h->d_queue;
  ta.ocache = h->ocache;
  ta.tot = h->d_tot;
  ta.stats = h->d_stats;
}

static int launch_team(mpcb_handle *h, const TeamArgs &ta) {
  const int grid = (int)std::min<int64_t>(h->B, (int64_t)h->num_sms * h->team_ctas);
  CK(cudaMemsetAsync(h->d_queue, 0, sizeof(int), h->stream));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h->timing) {
    while (h->ev_pool.size() < 2) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      h->ev_pool.push_back(e);
    }
    e0 = h->ev_pool[0];
    e1 = h->ev_pool[1];
    CK(cudaEventRecord(e0, h->stream));
  }
  ((team_fn)h->team_fn_ptr)<<<grid, h->team_threads, h->team_smem, h->stream>>>(ta);
  CK(cudaGetLastError());
  if (h->timing) {
    CK(cudaEventRecord(e1, h->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    h->ctr.admm_ms += ms;
  }
  h->ctr.kernel_launches += 1;
  h->ctr.admm_launches += 1;
  h->ctr.rounds += 1;
  return MPCB_OK;
}



// ------------------------------------------------------------------------------------------
// Tile kernel tables (tile.cuh): V padded for the DMMA fragment patterns, byte-indexed ELL.
static const int TILE_N = 81, TILE_M = 136, TILE_WARPS = 4, TILE_WA = 8, TILE_WAT = 11, TILE_WP = 4;
static int build_tile_tables(mpcb_handle *h) {
  const HostProblem &hp = h->hp;
  const mpcb_problem &p = hp.p;
  const int n = p.n, m = p.m;
  if (n != TILE_N || m != TILE_M) return MPCB_OK;
  const int ks = (n + 3) / 4, nt8 = (n + 7) / 8;
  const int ldv = ((4 * ks - 4 + 15) / 16) * 16 + 4, ldm = ((m - 4 + 15) / 16) * 16 + 4;
  const int nX = 4 * (p.Nx + 1);
  TileHdr hd;
  memset(&hd, 0, sizeof hd);
  std::vector<std::vector<unsigned char>> blobs(4);
  for (int v = 0; v < 4; ++v) {
    std::vector<double> A;
    variant_matrix(hp, v, A);
    int off = 0;
    auto take = [&](size_t bytes) {
      const int o = off;
      off = align16(off + (int)bytes);
      return o;
    };
    TileHdr hv;
    memset(&hv, 0, sizeof hv);
    hv.off_V = take((size_t)8 * nt8 * 8 * ldv);
    hv.off_lam = take(8 * ldv); hv.off_q = take(8 * ldv); hv.off_D = take(8 * ldv); hv.off_Dinv = take(8 * ldv);
    hv.off_E = take(8 * ldm); hv.off_Einv = take(8 * ldm); hv.off_lt = take(8 * ldm); hv.off_ut = take(8 * ldm);
    hv.off_Av = take((size_t)8 * TILE_WA * ldm); hv.off_ATv = take((size_t)8 * TILE_WAT * ldv); hv.off_Pv = take((size_t)8 * TILE_WP * ldv);
    hv.off_Ac = take((size_t)TILE_WA * ldm); hv.off_ATc = take((size_t)TILE_WAT * ldv); hv.off_Pc = take((size_t)TILE_WP * ldv);
    hv.off_An = take(ldm); hv.off_ATn = take(ldv); hv.off_Pn = take(ldv);
    hv.off_flags = take(ldm); hv.off_pcode = take(ldm);
    hv.total = off;
    if (v == 0) hd = hv;
    std::vector<unsigned char> &b = blobs[v];
    b.assign(hv.total, 0);
    double *Vd = reinterpret_cast<double *>(b.data() + hv.off_V);
    const double *V = hp.V.data() + (size_t)v * n * n;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) Vd[(size_t)i * ldv + j] = V[(size_t)i * n + j];
    auto dv = [&](int o) { return reinterpret_cast<double *>(b.data() + o); };
    for (int j = 0; j < n; ++j) {
      dv(hv.off_lam)[j] = hp.lam[(size_t)v * n + j];
      dv(hv.off_q)[j] = hp.q_s[j];
      dv(hv.off_D)[j] = hp.D[j];
      dv(hv.off_Dinv)[j] = 1.0 / hp.D[j];
    }
    for (int i = 0; i < m; ++i) {
      dv(hv.off_E)[i] = hp.E[i];
      dv(hv.off_Einv)[i] = 1.0 / hp.E[i];
      dv(hv.off_lt)[i] = hp.l_s[i];
      dv(hv.off_ut)[i] = hp.u_s[i];
      unsigned char f = 0;
      if (hp.l_s[i] < -1e30 * 1e-4) f |= 1;
      if (hp.u_s[i] > 1e30 * 1e-4) f |= 2;
      if (hp.ctype[i] == 1) f |= 4;
      if (hp.ctype[i] == -1) f |= 8;
      b[hv.off_flags + i] = f;
      unsigned char pc = 0;
      if (i < 4) pc = 1;
      else if (i == m - 2) pc = 3;
      else if (i == m - 1) pc = 4;
      else if (i >= nX && i < nX + 5 * (p.Nb + 1) && (i - nX) % 5 == 3) pc = 2;
      b[hv.off_pcode + i] = pc;
      int k = 0;
      for (int c = 0; c < n; ++c) {
        const double val = A[(size_t)i * n + c];
        if (val != 0.0) {
          if (k >= TILE_WA) return MPCB_OK;
          dv(hv.off_Av)[(size_t)k * ldm + i] = val;
          b[hv.off_Ac + (size_t)k * ldm + i] = (unsigned char)c;
          ++k;
        }
      }
      b[hv.off_An + i] = (unsigned char)k;
    }
    for (int j = 0; j < n; ++j) {
      int k = 0;
      for (int i = 0; i < m; ++i) {
        const double val = A[(size_t)i * n + j];
        if (val != 0.0) {
          if (k >= TILE_WAT) return MPCB_OK;
          dv(hv.off_ATv)[(size_t)k * ldv + j] = val;
          b[hv.off_ATc + (size_t)k * ldv + j] = (unsigned char)i;
          ++k;
        }
      }
      b[hv.off_ATn + j] = (unsigned char)k;
      k = 0;
      for (int c = 0; c < n; ++c) {
        const double val = hp.P_s[(size_t)j * n + c];
        if (val != 0.0) {
          if (k >= TILE_WP) return MPCB_OK;
          dv(hv.off_Pv)[(size_t)k * ldv + j] = val;
          b[hv.off_Pc + (size_t)k * ldv + j] = (unsigned char)c;
          ++k;
        }
      }
      b[hv.off_Pn + j] = (unsigned char)k;
    }
  }
  h->tile_warps = TILE_WARPS;
  h->tile_smem = (size_t)hd.total + (size_t)TILE_WARPS * 8 * (ldv + 3 * ldm) * 8;
  if (h->tile_smem > 227 * 1024) return MPCB_OK;
  for (int v = 0; v < 4; ++v) {
    CK(cudaMalloc(&h->d_tile_blob[v], hd.total));
    CK(cudaMemcpy(h->d_tile_blob[v], blobs[v].data(), hd.total, cudaMemcpyHostToDevice));
  }
  CK(cudaFuncSetAttribute((const void *)admm_tile_kernel<TILE_N, TILE_M, TILE_WARPS, TILE_WA, TILE_WAT, TILE_WP>,
                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->tile_smem));
  h->tile_hdr = hd;
  h->tile_ok = true;
  return MPCB_OK;
}


// ------------------------------------------------------------------------------------------
// Wave kernel tables (wave.cuh): Vp = D V in GEMM-position order per sign variant, [slot][quad member] scaling tables, and
// the unscaled block coefficients (Ad, Ad - Bd K, Bd, C, V_ecr, weights) the kernel uses as constants.  The coefficients
// are READ OFF the unscaled A / P the caller passed and A is then rebuilt from them: any mismatch (another problem
// structure) leaves wave_ok false and the other solver blocks take the batch.
static int build_wave_tables(mpcb_handle *h) {
  const HostProblem &hp = h->hp;
  const mpcb_problem &p = hp.p;
  const int n = p.n, m = p.m, NX = WAVE_NX, NC = WAVE_NC, NB = WAVE_NB;
  if (n != WAVE_N || m != WAVE_M || p.Nx != NX || p.Nc != NC || p.Nb != NB || hp.A_u.empty() || getenv("MPCB_NO_WAVE")) return MPCB_OK;
  const int nX = 4 * (NX + 1), RB = nX + 5 * (NX + 1);
  const std::vector<double> &A = hp.A_u, &P = hp.P_u;
  auto Aat = [&](int r, int c) { return A[(size_t)r * n + c]; };
  auto Pat = [&](int r, int c) { return P[(size_t)r * n + c]; };
  WaveConst K;
  memset(&K, 0, sizeof K);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      K.Ad[4 * i + j] = Aat(4 + i, j);
      K.Acl[4 * i + j] = Aat(4 * (NC + 1) + i, 4 * NC + j);
      K.Q[4 * i + j] = Pat(i, j);
      K.QN[4 * i + j] = Pat(4 * NX + i, 4 * NX + j);
    }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 2; ++j) K.Bd[2 * i + j] = Aat(4 + i, nX + j);
  for (int i = 0; i < 5; ++i) {
    for (int j = 0; j < 4; ++j) K.C[4 * i + j] = Aat(nX + i, j);
    K.Vecr[i] = Aat(nX + i, nX + 2 + i);
    K.Rs[i] = Pat(nX + 2 + i, nX + 2 + i);
  }
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) K.Ru[2 * i + j] = Pat(nX + i, nX + j);
  for (int j = 0; j < 4; ++j) {
    K.qx[j] = p.c * hp.q_u[j];
    K.qN[j] = p.c * hp.q_u[4 * NX + j];
  }
  K.ulim[0] = hp.u_u[RB];
  K.ulim[1] = hp.u_u[RB + 1];
  K.r_p = hp.l_u[nX + 2];
  // ---- rebuild A, P, q, l, u from the coefficients and compare
  {
    std::vector<double> A2((size_t)m * n, 0.0), P2((size_t)n * n, 0.0), q2(n, 0.0);
    for (int r = 0; r < nX; ++r) A2[(size_t)r * n + r] = -1.0;
    for (int k = 1; k <= NX; ++k)
      for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) A2[(size_t)(4 * k + i) * n + 4 * (k - 1) + j] += (k <= NC) ? K.Ad[4 * i + j] : K.Acl[4 * i + j];
        if (k <= NC)
          for (int j = 0; j < 2; ++j) A2[(size_t)(4 * k + i) * n + nX + 7 * (k - 1) + j] = K.Bd[2 * i + j];
        if (i < 2) A2[(size_t)(4 * k + i) * n + n - 2 + i] = 1.0;
      }
    for (int k = 0; k <= NX; ++k)
      for (int i = 0; i < 5; ++i) {
        for (int j = 0; j < 4; ++j) A2[(size_t)(nX + 5 * k + i) * n + 4 * k + j] = (i < 3) ? K.C[4 * i + j] : (i == 3 ? (j >= 2 ? 1.0 : 0.0) : (j == 1 ? 1.0 : 0.0));
        if (k < NC) A2[(size_t)(nX + 5 * k + i) * n + nX + 7 * k + 2 + i] = K.Vecr[i];
      }
    for (int q = 0; q < 7 * NC; ++q) A2[(size_t)(RB + q) * n + nX + q] = 1.0;
    A2[(size_t)(m - 2) * n + n - 2] = A2[(size_t)(m - 1) * n + n - 1] = 1.0;
    for (int k = 0; k <= NX; ++k)
      for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) P2[(size_t)(4 * k + i) * n + 4 * k + j] = (k < NX) ? K.Q[4 * i + j] : K.QN[4 * i + j];
        q2[4 * k + i] = ((k < NX) ? K.qx[i] : K.qN[i]);
      }
    for (int k = 0; k < NC; ++k) {
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) P2[(size_t)(nX + 7 * k + i) * n + nX + 7 * k + j] = K.Ru[2 * i + j];
      for (int i = 0; i < 5; ++i) P2[(size_t)(nX + 7 * k + 2 + i) * n + nX + 7 * k + 2 + i] = K.Rs[i];
    }
    P2[(size_t)(n - 2) * n + n - 2] = P2[(size_t)(n - 1) * n + n - 1] = 1.0;
    for (size_t q = 0; q < A2.size(); ++q)
      if (A2[q] != A[q]) return MPCB_OK;
    for (size_t q = 0; q < P2.size(); ++q)
      if (P2[q] != P[q]) return MPCB_OK;
    for (int j = 0; j < n; ++j)
      if (q2[j] != p.c * hp.q_u[j]) return MPCB_OK;
    // bounds the kernel hard-codes
    for (int k = 0; k <= NX; ++k)
      for (int i = 0; i < 5; ++i) {
        const double lo = hp.l_u[nX + 5 * k + i], hi = hp.u_u[nX + 5 * k + i];
        const double elo = (k <= NB) ? (i < 2 ? 1.0 : (i == 2 ? K.r_p : (i == 3 ? 0.0 : -INFINITY))) : -INFINITY;
        if (lo != elo) return MPCB_OK;
        if (!(k <= NB && i == 3) && hi != INFINITY) return MPCB_OK;
      }
    for (int k = 0; k < NC; ++k)
      for (int i = 0; i < 7; ++i) {
        const double lo = hp.l_u[RB + 7 * k + i], hi = hp.u_u[RB + 7 * k + i];
        if (i < 2 ? (lo != -K.ulim[i] || hi != K.ulim[i]) : (lo != 0.0 || hi != INFINITY)) return MPCB_OK;
      }
    for (int r = 4; r < nX; ++r)
      if (hp.l_u[r] != 0.0 || hp.u_u[r] != 0.0) return MPCB_OK;
  }
  K.c = p.c; K.cinv = 1.0 / p.c; K.sigma = p.sigma; K.alpha = p.alpha; K.eps_abs = p.eps_abs; K.eps_rel = p.eps_rel;
  K.eps_pinf = p.eps_prim_inf; K.adapt_tol = p.adaptive_rho_tolerance; K.qn_unscaled = h->qn_unscaled; K.qn_scaled = h->qn_scaled;
  K.check_every = p.check_termination; K.adaptive = p.adaptive_rho; K.adapt_interval = std::max(1, p.adaptive_rho_interval);
  K.max_iter = p.max_iter;
  // ---- ownership maps: (slot, quad member) -> natural variable / row index, variable -> GEMM position
  auto var_of = [&](int slot, int c) -> int {
    if (slot < 12) { const int r = slot / 4, j = slot % 4, k = 4 * r + c; return k <= NX ? 4 * k + j : -1; }
    if (slot < 16) { const int r = (slot - 12) / 2, j = (slot - 12) % 2, k = 4 * r + c; return (k >= 1 && k <= NC) ? nX + 7 * (k - 1) + j : -1; }
    if (slot < 26) { const int r = (slot - 16) / 5, j = (slot - 16) % 5, k = 4 * r + c; return k < NC ? nX + 7 * k + 2 + j : -1; }
    return n - 2 + (slot - 26);
  };
  auto row_of = [&](int slot, int c) -> int {
    if (slot < 12) { const int r = slot / 4, i = slot % 4, k = 4 * r + c; return k <= NX ? 4 * k + i : -1; }
    if (slot < 27) { const int r = (slot - 12) / 5, i = (slot - 12) % 5, k = 4 * r + c; return k <= NX ? nX + 5 * k + i : -1; }
    if (slot < 31) { const int r = (slot - 27) / 2, i = (slot - 27) % 2, k = 4 * r + c; return (k >= 1 && k <= NC) ? RB + 7 * (k - 1) + i : -1; }
    if (slot < 41) { const int r = (slot - 31) / 5, i = (slot - 31) % 5, k = 4 * r + c; return k < NC ? RB + 7 * k + 2 + i : -1; }
    return m - 2 + (slot - 41);
  };
  std::vector<int> pos_of(n, -1);
  for (int c = 0; c < 4; ++c) {
    for (int r = 0; r < 3; ++r)
      for (int j = 0; j < 4; ++j) { const int v = var_of(WVS_X(r, j), c); if (v >= 0) pos_of[v] = 16 * r + 4 * j + c; }
    for (int r = 0; r < 2; ++r) {
      for (int j = 0; j < 2; ++j) { const int v = var_of(WVS_U(r, j), c); if (v >= 0) pos_of[v] = (r == 0) ? 48 + 3 * j + (c - 1) : 54 + 2 * j + c; }
      for (int j = 0; j < 5; ++j) { const int v = var_of(WVS_S(r, j), c); if (v >= 0) pos_of[v] = (r == 0) ? 58 + 4 * j + c : 78 + j; }
    }
  }
  pos_of[n - 2] = 35;
  pos_of[n - 1] = 39;
  {
    std::vector<int> seen(WAVE_LD, 0);
    for (int v = 0; v < n; ++v) {
      if (pos_of[v] < 0 || pos_of[v] >= WAVE_LD || seen[pos_of[v]]++) return fail(MPCB_ERR_INVALID, "wave kernel: GEMM position map is not a bijection");
    }
  }
  int off = 0;
  auto take = [&](size_t bytes) {
    const int o = off;
    off = align16(off + (int)bytes);
    return o;
  };
  WaveHdr hd;
  memset(&hd, 0, sizeof hd);
  hd.off_Vp = take((size_t)8 * (88 * WAVE_LD + 8));
  hd.off_lam = take(8 * 88);
  hd.off_sgD = take(8 * WAVE_NVS * 4); hd.off_Dv = take(8 * WAVE_NVS * 4); hd.off_Dinv = take(8 * WAVE_NVS * 4);
  hd.off_ka = take(8 * WAVE_NRS * 4); hd.off_kb = take(8 * WAVE_NRS * 4); hd.off_kie = take(8 * WAVE_NRS * 4);
  hd.off_Ev = take(8 * WAVE_NRS * 4); hd.off_Einv = take(8 * WAVE_NRS * 4);
  hd.off_M1 = take(8 * 32);
  hd.total = off;
  h->wave_smem = (size_t)hd.total + (size_t)WAVE_WARPS * 8 * (WAVE_LD + WAVE_LDV + WAVE_LD) * 8;
  h->wave_smem4 = (size_t)hd.total + (size_t)4 * 8 * (WAVE_LD + WAVE_LDV + WAVE_LD) * 8;
  if (h->wave_smem > 227 * 1024) return MPCB_OK;
  for (int v = 0; v < 4; ++v) {
    std::vector<unsigned char> b(hd.total, 0);
    auto dv = [&](int o) { return reinterpret_cast<double *>(b.data() + o); };
    const double *V = hp.V.data() + (size_t)v * n * n;
    for (int var = 0; var < n; ++var)
      for (int J = 0; J < n; ++J) dv(hd.off_Vp)[(size_t)pos_of[var] * WAVE_LD + J] = hp.D[var] * V[(size_t)var * n + J];
    for (int J = 0; J < n; ++J) dv(hd.off_lam)[J] = hp.lam[(size_t)v * n + J];
    for (int slot = 0; slot < WAVE_NVS; ++slot)
      for (int c = 0; c < 4; ++c) {
        const int var = var_of(slot, c);
        const double D = var >= 0 ? hp.D[var] : 1.0;
        dv(hd.off_sgD)[4 * slot + c] = var >= 0 ? p.sigma / (D * D) : 0.0;
        dv(hd.off_Dv)[4 * slot + c] = D;
        dv(hd.off_Dinv)[4 * slot + c] = 1.0 / D;
      }
    for (int slot = 0; slot < WAVE_NRS; ++slot)
      for (int c = 0; c < 4; ++c) {
        const int row = row_of(slot, c);
        const double E = row >= 0 ? hp.E[row] : 1.0;
        // kap_i = ka * rho + kb: free rows carry the constant rho_min E^2, the others the class factor (1e3 on equalities) x E^2
        double ka = 1.0, kb = 0.0;
        if (row >= 0) {
          if (hp.ctype[row] == -1) { ka = 0.0; kb = MPCB_RHO_MIN * E * E; }
          else ka = (hp.ctype[row] == 1 ? MPCB_RHO_EQ : 1.0) * E * E;
        }
        dv(hd.off_ka)[4 * slot + c] = ka;
        dv(hd.off_kb)[4 * slot + c] = kb;
        dv(hd.off_kie)[4 * slot + c] = ka != 0.0 ? 1.0 / ka : 1.0;       // free rows keep y = 0: their 1 / kap is never used
        dv(hd.off_Ev)[4 * slot + c] = E;
        dv(hd.off_Einv)[4 * slot + c] = 1.0 / E;
      }
    for (int q = 0; q < 16; ++q) {
      dv(hd.off_M1)[q] = K.Ad[q];
      dv(hd.off_M1)[16 + q] = K.Acl[q];
    }
    CK(cudaMalloc(&h->d_wave_blob[v], hd.total));
    CK(cudaMemcpy(h->d_wave_blob[v], b.data(), hd.total, cudaMemcpyHostToDevice));
  }
  CK(cudaFuncSetAttribute((const void *)admm_wave_kernel<WAVE_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->wave_smem));
  CK(cudaFuncSetAttribute((const void *)admm_wave_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->wave_smem4));
  h->wave_hdr = hd;
  h->wave_k = K;
  h->wave_ok = true;
  if (const char *e = getenv("MPCB_WAVE_MIN_LANES")) h->wave_min_lanes = atoll(e);
  return MPCB_OK;
}

// ------------------------------------------------------------------------------------------
// Per-lane path tables: sparsity pattern of A (CSR + CSC view) with entry kinds, P in COO, unscaled vectors.
template <typename Tv>
static int gen_upload(mpcb_handle *h, const std::vector<Tv> &v, const Tv **out) {
  void *d = nullptr;
  CK(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(Tv)));
  CK(cudaMemcpy(d, v.data(), v.size() * sizeof(Tv), cudaMemcpyHostToDevice));
  h->gen_bufs.push_back(d);
  *out = (const Tv *)d;
  return MPCB_OK;
}

static int build_generic_tables(mpcb_handle *h) {
  int g0_nlong = 0, g0_longcols[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const HostProblem &hp = h->hp;
  const mpcb_problem &p = hp.p;
  const int n = p.n, m = p.m, nX = 4 * (p.Nx + 1);
  std::vector<int> rowptr(m + 1, 0), colidx, colptr(n + 1, 0), rowidx, cscpos, prow, pcol;
  std::vector<double> base, pval;
  std::vector<unsigned char> kind;
  for (int r = 0; r < m; ++r) {
    for (int c = 0; c < n; ++c) {
      double v = hp.A_u[(size_t)r * n + c];
      unsigned char k = 0;
      if (r >= nX && r < nX + 5 * (p.Nx + 1)) {
        const int blk = (r - nX) / 5, jj = (r - nX) % 5;
        if (jj == 3 && c == 4 * blk + 2) { k = 1; v = fabs(v); }
        else if (jj == 3 && c == 4 * blk + 3) { k = 2; v = fabs(v); }
        else if (jj == 4 && c == 4 * blk) { k = 3; v = -1.0; }              // C[4,0] = -slope (simhelpers.py:108)
      }
      if (v != 0.0) {
        colidx.push_back(c);
        base.push_back(v);
        kind.push_back(k);
      }
    }
    rowptr[r + 1] = (int)colidx.size();
  }
  const int nnz = (int)colidx.size();
  for (int e = 0; e < nnz; ++e) colptr[colidx[e] + 1]++;
  for (int c = 0; c < n; ++c) colptr[c + 1] += colptr[c];
  rowidx.resize(nnz);
  cscpos.resize(nnz);
  {
    std::vector<int> fill(colptr.begin(), colptr.end() - 1);
    for (int r = 0; r < m; ++r)
      for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) {
        const int q = fill[colidx[e]]++;
        rowidx[q] = r;
        cscpos[q] = e;
      }
  }
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c)
      if (hp.P_u[(size_t)r * n + c] != 0.0) {
        prow.push_back(r);
        pcol.push_back(c);
        pval.push_back(hp.P_u[(size_t)r * n + c]);
      }
  g0_nlong = 0;
  for (int c = 0; c < n; ++c)
    if (colptr[c + 1] - colptr[c] > 32) {
      if (g0_nlong == 8) return fail(MPCB_ERR_INVALID, "per-lane path: more than 8 columns of A with over 32 entries");
      g0_longcols[g0_nlong++] = c;
    }
  std::vector<int> prowptr(n + 1, 0);
  for (int r : prow) prowptr[r + 1]++;
  for (int r = 0; r < n; ++r) prowptr[r + 1] += prowptr[r];
  GenArgs &g = h->gproto;
  memset(&g, 0, sizeof g);
  g.n = n; g.m = m; g.nX = nX; g.Nx = p.Nx; g.Nb = p.Nb; g.Nc = p.Nc; g.uoff = nX;
  g.nnzA = nnz; g.nnzP = (int)pval.size(); g.scaling = p.scaling;
  g.nlong = g0_nlong;
  memcpy(g.longcols, g0_longcols, sizeof g.longcols);
  RC(gen_upload(h, rowptr, &g.rowptr)); RC(gen_upload(h, colidx, &g.colidx)); RC(gen_upload(h, colptr, &g.colptr));
  RC(gen_upload(h, rowidx, &g.rowidx)); RC(gen_upload(h, cscpos, &g.cscpos)); RC(gen_upload(h, base, &g.baseA));
  RC(gen_upload(h, kind, &g.kindA)); RC(gen_upload(h, prow, &g.prow)); RC(gen_upload(h, pcol, &g.pcol));
  RC(gen_upload(h, prowptr, &g.prowptr));
  RC(gen_upload(h, pval, &g.pval)); RC(gen_upload(h, hp.q_u, &g.q_u)); RC(gen_upload(h, hp.l_u, &g.l_u));
  RC(gen_upload(h, hp.u_u, &g.u_u));
  g.sigma = p.sigma; g.alpha = p.alpha; g.eps_abs = p.eps_abs; g.eps_rel = p.eps_rel; g.eps_pinf = p.eps_prim_inf;
  g.adapt_tol = p.adaptive_rho_tolerance;
  g.rho0 = std::min(std::max(p.rho0, MPCB_RHO_MIN), MPCB_RHO_MAX);
  g.check_every = p.check_termination; g.adaptive = p.adaptive_rho; g.adapt_interval = std::max(1, p.adaptive_rho_interval);
  g.max_iter = p.max_iter;
  g.sc = h->sc;
  g.has_debris = p.has_debris;
  g.dcx = p.debris_center[0]; g.dcy = p.debris_center[1]; g.dside = p.debris_side; g.ddetect = p.debris_detect;
  memcpy(g.verts, p.debris_verts, sizeof g.verts);
  memcpy(g.Kd, p.K_dead, sizeof g.Kd);
  memcpy(g.Kid, p.Ki_dead, sizeof g.Kid);
  auto ev = [](int c) { return (size_t)((c + 1) & ~1); };
  // + the operator's shared-memory part (generic.cuh): two pivot-row buffers and the row entries beyond GEN_TMD
  const int npad = (n + 3) & ~3, tmd = std::min(npad, GEN_TMD), TS = (n + 31) & ~31;
  h->gen_smem = 8 * (7 * ev(n) + 10 * ev(m) + ev(nnz) + ev(g.nnzP) + ev(16 * (GEN_THREADS / 32)) + ev(n) + ev((m + 1) / 2 + 1) +
                     ev(2 * npad) + ev((npad - tmd) * TS) + 8) + sizeof(GenLane) + 64 +
                2 * (size_t)(((m + 2) & ~1) + ((n + 2) & ~1) + 3 * ((nnz + 1) & ~1)) + 16;      // + the uint16 copies of the sparsity pattern
  if (nnz > 65535 || m > 65535) return fail(MPCB_ERR_INVALID, "problem too large for the per-lane path (16-bit sparsity indices)");
  if (n > GEN_THREADS) return fail(MPCB_ERR_INVALID, "problem too large for the per-lane path (one thread per row of the operator)");
  if (h->gen_smem > 227 * 1024) return fail(MPCB_ERR_INVALID, "problem too large for the per-lane path");
  CK(cudaFuncSetAttribute((const void *)generic_lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->gen_smem));
  h->gen_grid = h->num_sms;              // one CTA per SM: it owns the SM's tensor memory
  h->generic_ok = true;
  return MPCB_OK;
}

static int launch_generic(mpcb_handle *h, GenArgs &g) {
  g.B = (int)h->B;
  g.xs = h->xs; g.zs = h->zs; g.ys = h->ys; g.rho = h->rho; g.u0 = h->u0; g.iter = h->iter; g.status = h->status;
  g.queue = h->d_queue;
  g.tot = h->d_tot;
  g.stats = h->d_stats;
  const int grid = (int)std::min<int64_t>(h->B, h->gen_grid);
  CK(cudaMemsetAsync(h->d_queue, 0, sizeof(int), h->stream));
  if (h->timing) {
    while (h->ev_pool.size() < 2) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      h->ev_pool.push_back(e);
    }
    CK(cudaEventRecord(h->ev_pool[0], h->stream));
  }
  generic_lane_kernel<<<grid, GEN_THREADS, h->gen_smem, h->stream>>>(g);
  CK(cudaGetLastError());
  if (h->timing) {
    CK(cudaEventRecord(h->ev_pool[1], h->stream));
    CK(cudaEventSynchronize(h->ev_pool[1]));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev_pool[0], h->ev_pool[1]));
    h->ctr.admm_ms += ms;
  }
  h->ctr.kernel_launches += 1;
  h->ctr.admm_launches += 1;
  h->ctr.rounds += 1;
  return MPCB_OK;
}

// ------------------------------------------------------------------------------------------
extern "C" int mpcb_abi_version(void) { return MPCB_ABI_VERSION; }
extern "C" const char *mpcb_last_error(void) { return g_err.c_str(); }

static int create_tail(mpcb_handle *h, bool debris) {
  if (!debris) CK(cudaFuncSetAttribute((const void *)h->kern.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CK(cudaMalloc(&h->d_tot, 16 * sizeof(unsigned long long)));
  CK(cudaMemset(h->d_tot, 0, 16 * sizeof(unsigned long long)));
  CK(cudaMalloc(&h->d_queue, sizeof(int)));
  CK(cudaMalloc(&h->d_stats, 2 * MPCB_NSTATS * sizeof(double)));
  CK(cudaMallocHost(&h->h_cnt, 8 * sizeof(int)));
  CK(cudaEventCreate(&h->ev_t0));
  CK(cudaEventCreate(&h->ev_t1));
  CK(cudaEventCreateWithFlags(&h->ev_wait, cudaEventDisableTiming));
  return MPCB_OK;
}

extern "C" int mpcb_create(const mpcb_problem *pr, int device, mpcb_handle **out) {
  if (!pr || !out) return fail(MPCB_ERR_INVALID, "null argument");
  *out = nullptr;
  const int n = pr->n, m = pr->m;
  if (pr->Nx < 1 || pr->Nc < 1 || pr->Nc > pr->Nx || pr->Nb < 0 || pr->Nb > pr->Nx)
    return fail(MPCB_ERR_INVALID, "bad horizons");
  if (pr->estimator != MPCB_EST_UKF && pr->estimator != MPCB_EST_KF) return fail(MPCB_ERR_INVALID, "unknown estimator");
  if (n != 4 * (pr->Nx + 1) + 7 * pr->Nc + 2 || m != 9 * (pr->Nx + 1) + 7 * pr->Nc + 2)
    return fail(MPCB_ERR_INVALID, "n/m do not match the horizons");
  const bool debris = pr->has_debris != 0;
  if (debris) {
    if (!pr->P_u || !pr->q_u || !pr->A_u || !pr->l_u || !pr->u_u)
      return fail(MPCB_ERR_INVALID, "debris problems need the unscaled P_u, q_u, A_u, l_u, u_u");
  } else if (!pr->P_s || !pr->q_s || !pr->A_s || !pr->l_s || !pr->u_s || !pr->D || !pr->E || !pr->ctype || !pr->V || !pr->lam) {
    return fail(MPCB_ERR_INVALID, "null table pointer");
  }
  if (pr->check_termination < 1 || pr->max_iter % pr->check_termination != 0 ||
      (pr->adaptive_rho && (pr->adaptive_rho_interval < 1 || pr->adaptive_rho_interval % pr->check_termination != 0)))
    return fail(MPCB_ERR_INVALID, "max_iter and adaptive_rho_interval must be multiples of check_termination");
  if (pr->max_iter > INT16_MAX) return fail(MPCB_ERR_INVALID, "max_iter above 32767: the per-solve iteration telemetry is int16");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(MPCB_ERR_INVALID, "no such CUDA device");
  CK(cudaSetDevice(device));
  mpcb_handle *h = new mpcb_handle();
  h->device = device;
  memset(&h->ctr, 0, sizeof h->ctr);
  memset(&h->ls, 0, sizeof h->ls);
  HostProblem &hp = h->hp;
  hp.p = *pr;
  if (debris) {
    hp.P_u.assign(pr->P_u, pr->P_u + (size_t)n * n);
    hp.q_u.assign(pr->q_u, pr->q_u + n);
    hp.A_u.assign(pr->A_u, pr->A_u + (size_t)m * n);
    hp.l_u.assign(pr->l_u, pr->l_u + m);
    hp.u_u.assign(pr->u_u, pr->u_u + m);
  } else {
    hp.P_s.assign(pr->P_s, pr->P_s + (size_t)n * n);
    hp.q_s.assign(pr->q_s, pr->q_s + n);
    hp.A_s.assign(pr->A_s, pr->A_s + (size_t)m * n);
    hp.l_s.assign(pr->l_s, pr->l_s + m);
    hp.u_s.assign(pr->u_s, pr->u_s + m);
    hp.D.assign(pr->D, pr->D + n);
    hp.E.assign(pr->E, pr->E + m);
    hp.V.assign(pr->V, pr->V + (size_t)4 * n * n);
    hp.lam.assign(pr->lam, pr->lam + (size_t)4 * n);
    hp.ctype.assign(pr->ctype, pr->ctype + m);
    if (pr->P_u && pr->q_u && pr->A_u && pr->l_u && pr->u_u) {     // unscaled data: the wave kernel iterates in unscaled variables
      hp.P_u.assign(pr->P_u, pr->P_u + (size_t)n * n);
      hp.q_u.assign(pr->q_u, pr->q_u + n);
      hp.A_u.assign(pr->A_u, pr->A_u + (size_t)m * n);
      hp.l_u.assign(pr->l_u, pr->l_u + m);
      hp.u_u.assign(pr->u_u, pr->u_u + m);
    }
  }
  hp.p.P_s = hp.p.q_s = hp.p.A_s = hp.p.l_s = hp.p.u_s = hp.p.D = hp.p.E = hp.p.V = hp.p.lam = nullptr;
  hp.p.ctype = nullptr;
  hp.p.P_u = hp.p.q_u = hp.p.A_u = hp.p.l_u = hp.p.u_u = nullptr;
  SimConst &sc = h->sc;
  memcpy(sc.Ad, pr->Ad, sizeof sc.Ad);
  memcpy(sc.Bd, pr->Bd, sizeof sc.Bd);
  memcpy(sc.Ao, pr->Ao, sizeof sc.Ao);
  memcpy(sc.Bou, pr->Bou, sizeof sc.Bou);
  memcpy(sc.Qw, pr->Qw, sizeof sc.Qw);
  memcpy(sc.Kpf, pr->Kpf, sizeof sc.Kpf);
  memcpy(sc.Kif, pr->Kif, sizeof sc.Kif);
  memcpy(sc.xr, pr->xr, sizeof sc.xr);
  sc.umax0 = pr->umax0; sc.r_p = pr->r_p; sc.r_tol = pr->r_tol; sc.suc_dist = pr->suc_dist;
  sc.suc_ang_deg = pr->suc_ang_deg; sc.mean_mtn = pr->mean_mtn;
  sc.in_track = pr->in_track; sc.delta_v = pr->delta_v; sc.is_reject = pr->is_reject; sc.has_noise = pr->has_noise;
  sc.noise_length = std::max(1, pr->noise_length);
  sc.estimator = pr->estimator;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  h->num_sms = prop.multiProcessorCount;
  int rc = MPCB_OK;
  if (debris) {
    rc = build_generic_tables(h);
  } else {
    rc = build_tables(h);
    if (rc == MPCB_OK) {
      set_warps(h, h->warps);
      rc = build_team_tables(h);
      if (rc == MPCB_OK) rc = build_tile_tables(h);
      if (rc == MPCB_OK) rc = build_wave_tables(h);
    }
  }
  if (rc != MPCB_OK) {
    mpcb_destroy(h);
    return rc;
  }
  rc = create_tail(h, debris);       // a failure past this point still owns h: release it through mpcb_destroy
  if (rc != MPCB_OK) {
    const std::string keep = g_err;
    mpcb_destroy(h);
    g_err = keep;
    return rc;
  }
  *out = h;
  return MPCB_OK;
}

static void free_batch(mpcb_handle *h) {
  cudaFree(h->scache); cudaFree(h->scache_rho); cudaFree(h->scache_var);
  h->scache = h->scache_rho = nullptr; h->scache_var = nullptr; h->scache_tried = false;
  cudaFree(h->xs); cudaFree(h->zs); cudaFree(h->ys); cudaFree(h->rho); cudaFree(h->par); cudaFree(h->u0);
  cudaFree(h->iter); cudaFree(h->status); cudaFree(h->flip); cudaFree(h->lane_state); cudaFree(h->cnt);
  cudaFree(h->list); cudaFree(h->lane_f64); cudaFree(h->lane_i32); cudaFree(h->lane_fd);
  h->lane_fd = nullptr;
  h->xs = h->zs = h->ys = h->rho = h->par = h->u0 = h->lane_f64 = nullptr;
  h->iter = h->status = h->flip = h->cnt = h->list = h->lane_i32 = nullptr;
  h->lane_state = nullptr;
  h->B = 0;
}

extern "C" int mpcb_destroy(mpcb_handle *h) {
  if (!h) return MPCB_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  free_batch(h);
  for (int v = 0; v < 4; ++v) {
    cudaFree(h->d_blob[v]);
    cudaFree(h->d_V[v]);
  }
  cudaFree(h->d_tot);
  for (int v = 0; v < 4; ++v) cudaFree(h->d_tile_blob[v]);
  for (int v = 0; v < 4; ++v) cudaFree(h->d_wave_blob[v]);
  for (void *q : h->gen_bufs) cudaFree(q);
  for (StageBuf &b : h->stage_pool) cudaFree(b.ptr);
  cudaFree(h->d_tblob);
  cudaFree(h->d_lam);
  cudaFree(h->d_queue);
  for (int v = 0; v < 4; ++v) cudaFree(h->d_Vk[v]);
  cudaFree(h->ocache);
  cudaFree(h->d_stats);
  if (h->h_cnt) cudaFreeHost(h->h_cnt);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  if (h->ev_t0) cudaEventDestroy(h->ev_t0);
  if (h->ev_t1) cudaEventDestroy(h->ev_t1);
  if (h->ev_wait) cudaEventDestroy(h->ev_wait);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return MPCB_OK;
}

static int reset_solver_state(mpcb_handle *h) {
  const int64_t B = h->B;
  const int n = h->hp.p.n, m = h->hp.p.m;
  CK(cudaMemsetAsync(h->xs, 0, (size_t)B * n * 8, h->stream));
  CK(cudaMemsetAsync(h->zs, 0, (size_t)B * m * 8, h->stream));
  CK(cudaMemsetAsync(h->ys, 0, (size_t)B * m * 8, h->stream));
  CK(cudaMemsetAsync(h->flip, 0, (size_t)B * 4, h->stream));
  CK(cudaMemsetAsync(h->cnt, 0, 12 * sizeof(int), h->stream));
  CK(cudaMemsetAsync(h->d_tot, 0, 16 * sizeof(unsigned long long), h->stream));
  std::vector<double> r((size_t)B, std::min(std::max(h->hp.p.rho0, MPCB_RHO_MIN), MPCB_RHO_MAX));
  CK(cudaMemcpyAsync(h->rho, r.data(), (size_t)B * 8, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->sim_done = false;
  return MPCB_OK;
}

extern "C" int mpcb_batch_alloc(mpcb_handle *h, int64_t B) {
  if (!h || B < 1 || B > (int64_t)1 << 30) return fail(MPCB_ERR_INVALID, "bad batch size");
  CK(cudaSetDevice(h->device));
  if (h->B != B) {
    free_batch(h);
    const int n = h->hp.p.n, m = h->hp.p.m;
    CK(cudaMalloc(&h->xs, (size_t)B * n * 8));
    CK(cudaMalloc(&h->zs, (size_t)B * m * 8));
    CK(cudaMalloc(&h->ys, (size_t)B * m * 8));
    CK(cudaMalloc(&h->rho, (size_t)B * 8));
    CK(cudaMalloc(&h->par, (size_t)B * 7 * 8));
    CK(cudaMalloc(&h->u0, (size_t)B * 2 * 8));
    CK(cudaMalloc(&h->iter, (size_t)B * 4));
    CK(cudaMalloc(&h->status, (size_t)B * 4));
    CK(cudaMalloc(&h->flip, (size_t)B * 4));
    CK(cudaMalloc(&h->lane_state, (size_t)B));
    CK(cudaMalloc(&h->cnt, 12 * sizeof(int)));                       // two round buffers of 4 + the 4 deferred-lane counters
    CK(cudaMalloc(&h->list, (size_t)12 * B * sizeof(int)));
    // LaneSim: doubles 4+6+36+4+2+2+1+2+4 = 61 per lane, ints 7 per lane
    CK(cudaMalloc(&h->lane_f64, (size_t)B * 61 * 8));
    CK(cudaMalloc(&h->lane_i32, (size_t)B * 7 * 4));
    CK(cudaMalloc(&h->lane_fd, (size_t)B * 8));
    double *d = h->lane_f64;
    LaneSim &ls = h->ls;
    ls.xtrue = d; d += 4 * B;
    ls.ux = d; d += 6 * B;
    ls.uP = d; d += 36 * B;
    ls.xstore = d; d += 4 * B;
    ls.uprev = d; d += 2 * B;
    ls.unext = d; d += 2 * B;
    ls.xintf = d; d += B;
    ls.noise = d; d += 2 * B;
    ls.xfin = d; d += 4 * B;
    int *q = h->lane_i32;
    ls.step = q; q += B;
    ls.sub = q; q += B;
    ls.iterm = q; q += B;
    ls.succ = q; q += B;
    ls.nsolve = q; q += B;
    ls.variant = q; q += B;
    ls.ukf_clamp = q; q += B;
    h->B = B;
  }
  CK(cudaMemsetAsync(h->lane_f64, 0, (size_t)B * 61 * 8, h->stream));
  CK(cudaMemsetAsync(h->lane_i32, 0, (size_t)B * 7 * 4, h->stream));
  return reset_solver_state(h);
}

extern "C" int mpcb_set_timing(mpcb_handle *h, int enable) {
  if (!h) return fail(MPCB_ERR_INVALID, "null handle");
  h->timing = enable != 0;
  return MPCB_OK;
}

extern "C" int mpcb_get_counters(mpcb_handle *h, mpcb_counters *out) {
  if (!h || !out) return fail(MPCB_ERR_INVALID, "null argument");
  *out = h->ctr;
  return MPCB_OK;
}

extern "C" void *mpcb_stream(mpcb_handle *h) { return h ? (void *)h->stream : nullptr; }

extern "C" int mpcb_solver_blocks(mpcb_handle *h) {
  if (!h) return 0;
  return (h->kern.fn ? 1 : 0) | (h->team_ok ? 2 : 0) | (h->tile_ok ? 4 : 0) | (h->wave_ok ? 8 : 0) | (h->generic_ok ? 16 : 0);
}

extern "C" int mpcb_wait_stream(mpcb_handle *h, void *producer_stream) {
  if (!h) return fail(MPCB_ERR_INVALID, "null handle");
  CK(cudaSetDevice(h->device));
  CK(cudaEventRecord(h->ev_wait, (cudaStream_t)producer_stream));
  CK(cudaStreamWaitEvent(h->stream, h->ev_wait, 0));
  return MPCB_OK;
}

// ------------------------------------------------------------------------------------------
static void fill_args(mpcb_handle *h, AdmmArgs &aa, PostArgs &pa, int mode) {
  const mpcb_problem &p = h->hp.p;
  memset(&aa, 0, sizeof aa);
  aa.hdr = h->hdr;
  for (int v = 0; v < 4; ++v) {
    aa.blob[v] = h->d_blob[v];
    aa.Vg[v] = h->d_V[v];
  }
  aa.n = p.n; aa.m = p.m; aa.nX = 4 * (p.Nx + 1); aa.Nx = p.Nx; aa.Nb = p.Nb; aa.uoff = 4 * (p.Nx + 1);
  aa.B = (int)h->B;
  aa.sigma = p.sigma; aa.alpha = p.alpha; aa.eps_abs = p.eps_abs; aa.eps_rel = p.eps_rel; aa.eps_pinf = p.eps_prim_inf;
  aa.adapt_tol = p.adaptive_rho_tolerance; aa.cinv = 1.0 / p.c;
  aa.qn_unscaled = h->qn_unscaled; aa.qn_scaled = h->qn_scaled;
  aa.check_every = p.check_termination; aa.adaptive = p.adaptive_rho; aa.adapt_interval = std::max(1, p.adaptive_rho_interval);
  aa.max_iter = p.max_iter;
  aa.xs = h->xs; aa.zs = h->zs; aa.ys = h->ys; aa.rho = h->rho; aa.iter = h->iter; aa.status = h->status;
  aa.par = h->par; aa.u0 = h->u0; aa.lane_state = h->lane_state; aa.flip = h->flip; aa.iter_total = h->d_tot;
  memset(&pa, 0, sizeof pa);
  pa.sc = h->sc;
  pa.ls = h->ls;
  pa.mode = mode;
  pa.B = (int)h->B;
  pa.par = h->par; pa.u0 = h->u0; pa.rho = h->rho; pa.iter = h->iter; pa.status = h->status;
  pa.lane_state = h->lane_state;
  pa.solves_total = h->d_tot + 1;
}

// The lockstep round loop: every round runs one check_termination block of ADMM for all
// still-solving lanes, then the per-lane epilogue for the lanes whose solve ended.

// Which solver block runs this batch.  MPCB_SOLVER = tile | team | block forces one; the default is the persistent team
// kernel where an instantiation exists (n = 81), else the warp-per-lane block kernel.
static bool want_tile(const mpcb_handle *h) {
  if (!h->tile_ok) return false;
  const char *e = getenv("MPCB_SOLVER");
  if (e && *e) return strcmp(e, "tile") == 0;
  return h->B >= h->tile_min_lanes;
}
// 4-warp wave CTAs: between 3/4 of the SMs and all of them get one CTA of 32 lanes
static bool wave_small_window(const mpcb_handle *h) {
  if (getenv("MPCB_NO_WAVE4")) return false;
  const int64_t ctas = (h->B + 31) / 32;
  return 4 * ctas >= 3 * (int64_t)h->num_sms && ctas <= h->num_sms;
}
static bool want_wave(const mpcb_handle *h, bool round_based) {
  if (!h->wave_ok) return false;
  const char *e = getenv("MPCB_SOLVER");
  if (e && *e) return strcmp(e, "wave") == 0;
  // Large batches: rounds of the multi-RHS wave kernel (2x the team kernel's throughput once every SM holds 64 lanes), the
  // last lanes handed to the team kernel mid-flight.  Batches that give (nearly) every SM one 4-warp CTA of 32 lanes run the
  // 4-warp instantiation (measured on BASELINE config 2, 4096 lanes: 106 ms per step against 115 on the team kernel alone).
  // In between and below, the team kernel's whole-loop launch wins: rounds cost 25 iterations of wave latency each however
  // few lanes they carry.
  if (!round_based) return false;
  return h->B >= h->wave_min_lanes || wave_small_window(h);
}
static bool want_team(const mpcb_handle *h) {
  if (!h->team_ok) return false;
  const char *e = getenv("MPCB_SOLVER");
  if (e && *e) return strcmp(e, "team") == 0;
  return true;
}

static int run_rounds(mpcb_handle *h, AdmmArgs &aa, PostArgs &pa, int first_buf, bool use_tile = false, bool use_wave = false,
                      const TeamArgs *resume = nullptr, long resume_below = 0) {
  const int64_t B = h->B;
  // n = 81 family: each round solves its lanes to completion on the team kernel (list mode) instead of advancing them
  // check_termination iterations at a time; rounds then coincide with control steps
  const bool use_team = !use_tile && !use_wave && want_team(h);
  if (const char *e = getenv("MPCB_VISIT_ITERS")) h->visit_iters = std::max(-1, atoi(e));
  if (use_team && h->team_tm && !h->scache_tried) {
    h->scache_tried = true;
    const size_t per_lane = (size_t)(((h->hp.p.n + 3) / 4)) * (((2 * h->hp.p.n + 31) / 32) * 32) * 16;
    const char *e = getenv("MPCB_SCACHE_GB");
    const double cap = (e && *e) ? atof(e) : 24.0;
    if ((double)per_lane * (double)B <= cap * 1073741824.0) {
      if (cudaMalloc(&h->scache, per_lane * (size_t)B) == cudaSuccess && cudaMalloc(&h->scache_rho, (size_t)B * 8) == cudaSuccess &&
          cudaMalloc(&h->scache_var, (size_t)B * 4) == cudaSuccess) {
        CK(cudaMemsetAsync(h->scache_rho, 0xff, (size_t)B * 8, h->stream));      // NaN tags: every slot empty
        CK(cudaMemsetAsync(h->scache_var, 0xff, (size_t)B * 4, h->stream));
      } else {
        cudaGetLastError();
        cudaFree(h->scache); cudaFree(h->scache_rho); cudaFree(h->scache_var);
        h->scache = h->scache_rho = nullptr; h->scache_var = nullptr;
      }
    }
  }
  const int W = h->warps;
  const int pgrid = (int)((B + 127) / 128);
  int cur = first_buf;
  size_t ev_used = 0;
  while (true) {
    CK(cudaMemcpyAsync(h->h_cnt, h->cnt + 4 * cur, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    int grid = 0;
    long live = 0;
    for (int v = 0; v < 4; ++v) {
      grid += (h->h_cnt[v] + W - 1) / W;
      live += h->h_cnt[v];
    }
    if (live == 0) break;
    if (resume && live <= resume_below) {
      // few lanes left: a round would leave most SMs idle and still cost a full launch + 25-iteration latency.  The team kernel
      // takes the listed lanes over mid-flight (MODE_RESUME) and carries each to the end of its trajectory in ONE launch.
      TeamArgs ta = *resume;
      ta.cnt = h->cnt + 4 * cur;
      ta.list = h->list + (size_t)4 * B * cur;
      CK(cudaMemsetAsync(h->d_queue, 0, sizeof(int), h->stream));
      const int tgrid = (int)std::min<long>(live, (long)h->num_sms * h->team_ctas);
      if (h->timing) {
        while (h->ev_pool.size() < ev_used + 2) {
          cudaEvent_t e;
          CK(cudaEventCreate(&e));
          h->ev_pool.push_back(e);
        }
        CK(cudaEventRecord(h->ev_pool[ev_used++], h->stream));
      }
      ((team_fn)h->team_fn_ptr)<<<tgrid, h->team_threads, h->team_smem, h->stream>>>(ta);
      CK(cudaGetLastError());
      if (h->timing) CK(cudaEventRecord(h->ev_pool[ev_used++], h->stream));
      h->ctr.kernel_launches += 1;
      h->ctr.admm_launches += 1;
      h->ctr.rounds += 1;
      break;
    }
    aa.cnt = h->cnt + 4 * cur;
    aa.list = h->list + (size_t)4 * B * cur;
    pa.cnt_cur = h->cnt + 4 * cur;
    pa.cnt_next = h->cnt + 4 * (1 - cur);
    pa.list_next = h->list + (size_t)4 * B * (1 - cur);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->timing) {
      while (h->ev_pool.size() < ev_used + 2) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        h->ev_pool.push_back(e);
      }
      e0 = h->ev_pool[ev_used++];
      e1 = h->ev_pool[ev_used++];
      CK(cudaEventRecord(e0, h->stream));
    }
    if (use_team) {
      TeamArgs ta;
      fill_team_args(h, ta, MODE_QP_ONLY);
      ta.warm = 1;
      ta.cnt = aa.cnt; ta.list = aa.list; ta.par = aa.par; ta.lane_state = aa.lane_state;
      ta.scache = h->scache; ta.scache_rho = h->scache_rho; ta.scache_var = h->scache_var;
      // Visit budget: a round lasts as long as its longest visit, so cap a visit at about the work an average CTA has
      // this round (live/grid lanes x ~46 iterations); with a CTA per live lane nobody queues and solves run to completion.
      const long cta_max = (long)h->num_sms * h->team_ctas;
      if (h->visit_iters >= 0) ta.visit_iters = h->visit_iters;
      else if (live <= cta_max) ta.visit_iters = 0;
      else ta.visit_iters = (int)std::max<long>(4L * aa.check_every, (live / 10 / aa.check_every) * aa.check_every);
      CK(cudaMemsetAsync(h->d_queue, 0, sizeof(int), h->stream));
      const int tgrid = (int)std::min<long>(live, (long)h->num_sms * h->team_ctas);
      ((team_fn)h->team_fn_ptr)<<<tgrid, h->team_threads, h->team_smem, h->stream>>>(ta);
    } else if (use_wave) {
      WaveArgs wa;
      memset(&wa, 0, sizeof wa);
      wa.hdr = h->wave_hdr;
      wa.k = h->wave_k;
      for (int v = 0; v < 4; ++v) wa.blob[v] = h->d_wave_blob[v];
      wa.B = aa.B;
      wa.cnt = aa.cnt; wa.list = aa.list; wa.xs = aa.xs; wa.zs = aa.zs; wa.ys = aa.ys; wa.rho = aa.rho; wa.iter = aa.iter;
      wa.status = aa.status; wa.par = aa.par; wa.u0 = aa.u0; wa.lane_state = aa.lane_state; wa.flip = aa.flip;
      wa.iter_total = aa.iter_total;
      int wgrid = 0;
      const int ww = h->wave_warps_now;
      for (int v = 0; v < 4; ++v) wgrid += (h->h_cnt[v] + 8 * ww - 1) / (8 * ww);
      if (ww == 4) admm_wave_kernel<4><<<wgrid, 32 * 4, h->wave_smem4, h->stream>>>(wa);
      else admm_wave_kernel<WAVE_WARPS><<<wgrid, 32 * WAVE_WARPS, h->wave_smem, h->stream>>>(wa);
    } else if (use_tile) {
      TileArgs ta;
      memset(&ta, 0, sizeof ta);
      ta.hdr = h->tile_hdr;
      for (int v = 0; v < 4; ++v) ta.blob[v] = h->d_tile_blob[v];
      ta.n = aa.n; ta.m = aa.m; ta.uoff = aa.uoff; ta.B = aa.B;
      ta.sigma = aa.sigma; ta.alpha = aa.alpha; ta.eps_abs = aa.eps_abs; ta.eps_rel = aa.eps_rel; ta.eps_pinf = aa.eps_pinf;
      ta.adapt_tol = aa.adapt_tol; ta.cinv = aa.cinv; ta.qn_unscaled = aa.qn_unscaled; ta.qn_scaled = aa.qn_scaled;
      ta.check_every = aa.check_every; ta.adaptive = aa.adaptive; ta.adapt_interval = aa.adapt_interval; ta.max_iter = aa.max_iter;
      ta.cnt = aa.cnt; ta.list = aa.list; ta.xs = aa.xs; ta.zs = aa.zs; ta.ys = aa.ys; ta.rho = aa.rho; ta.iter = aa.iter;
      ta.status = aa.status; ta.par = aa.par; ta.u0 = aa.u0; ta.lane_state = aa.lane_state; ta.flip = aa.flip;
      ta.iter_total = aa.iter_total;
      int tgrid = 0;
      for (int v = 0; v < 4; ++v) tgrid += (h->h_cnt[v] + 8 * h->tile_warps - 1) / (8 * h->tile_warps);
      admm_tile_kernel<TILE_N, TILE_M, TILE_WARPS, TILE_WA, TILE_WAT, TILE_WP><<<tgrid, 32 * TILE_WARPS, h->tile_smem, h->stream>>>(ta);
    } else {
      h->kern.fn<<<grid, 32 * W, h->smem_bytes, h->stream>>>(aa);
    }
    if (h->timing) CK(cudaEventRecord(e1, h->stream));
    post_kernel<<<pgrid, 128, 0, h->stream>>>(pa);
    CK(cudaGetLastError());
    h->ctr.kernel_launches += 2;
    h->ctr.admm_launches += 1;
    h->ctr.rounds += 1;
    cur = 1 - cur;
  }
  if (h->timing) {
    CK(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i + 1 < ev_used; i += 2) {
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, h->ev_pool[i], h->ev_pool[i + 1]));
      h->ctr.admm_ms += ms;
    }
  }
  return MPCB_OK;
}

static int pull_totals(mpcb_handle *h) {
  unsigned long long t[16];
  CK(cudaMemcpyAsync(t, h->d_tot, sizeof t, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
#ifdef GEN_PROFILE
  if (t[10]) fprintf(stderr, "[generic profile] share of CTA time: values+scaling %.1f %% | assembly %.1f %% | inversion %.1f %% | iterations %.1f %% | "
                     "checks %.1f %% | rest of step %.1f %%   (%.2f Mcycles per solve, %.1f rebuilds per solve)\n", 100.0 * t[4] / t[10],
                     100.0 * t[5] / t[10], 100.0 * t[6] / t[10], 100.0 * t[7] / t[10], 100.0 * t[8] / t[10], 100.0 * t[9] / t[10],
                     t[1] ? (double)t[10] / t[1] / 1e6 : 0.0, t[1] ? (double)t[2] / t[1] : 0.0);
#endif
#ifdef TEAM_PROFILE
  fprintf(stderr, "[team profile] Mcycles: iterations %.1f checks %.1f rebuilds %.1f post %.1f setup %.1f | team total %.1f\n",
          t[4] / 1e6, t[5] / 1e6, t[6] / 1e6, t[7] / 1e6, t[8] / 1e6, t[9] / 1e6);
  if (t[1]) fprintf(stderr, "[team profile] post step, cycles per control step: select/clip/telemetry %.0f | UKF %.0f | refresh/telemetry %.0f\n",
                    (double)t[10] / t[1], (double)t[11] / t[1], (double)t[12] / t[1]);
  if (t[1]) fprintf(stderr, "[team profile] inside UKF: predict %.0f | cholesky %.0f | sigma+hx %.0f\n", (double)t[13] / t[1], (double)t[14] / t[1], (double)t[15] / t[1]);
#endif
  h->ctr.admm_iterations += (int64_t)t[0];
  h->ctr.operator_rebuilds += (int64_t)t[2];
  CK(cudaMemsetAsync(h->d_tot, 0, sizeof t, h->stream));
  return MPCB_OK;
}

// Staging of host <-> device batch arrays.  Device staging buffers are cached in the handle (grow-only,
// matched by request order), so repeated host-pointer calls do not pay cudaMalloc / cudaFree.
struct Stage {
  std::vector<StageBuf> *pool;
  size_t next = 0;
  explicit Stage(std::vector<StageBuf> *p) : pool(p) {}
  int grab(size_t bytes, void **out) {
    if (next == pool->size()) pool->push_back(StageBuf());
    StageBuf &b = (*pool)[next++];
    if (b.cap < bytes) {
      if (b.ptr) CK(cudaFree(b.ptr));
      b.ptr = nullptr;
      b.cap = 0;
      CK(cudaMalloc(&b.ptr, bytes));
      b.cap = bytes;
    }
    *out = b.ptr;
    return MPCB_OK;
  }
  template <typename T>
  int in(const T *src, size_t count, int on_device, cudaStream_t s, const T **out) {
    if (!src) { *out = nullptr; return MPCB_OK; }
    if (on_device) { *out = src; return MPCB_OK; }
    void *d = nullptr;
    int rc = grab(count * sizeof(T), &d);
    if (rc != MPCB_OK) return rc;
    CK(cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
    *out = (const T *)d;
    return MPCB_OK;
  }
  template <typename T>
  int out(T *dst, size_t count, int on_device, T **dev) {
    if (!dst) { *dev = nullptr; return MPCB_OK; }
    if (on_device) { *dev = dst; return MPCB_OK; }
    void *d = nullptr;
    int rc = grab(count * sizeof(T), &d);
    if (rc != MPCB_OK) return rc;
    *dev = (T *)d;
    return MPCB_OK;
  }
  template <typename T>
  int back(T *dst, const T *dev, size_t count, int on_device, cudaStream_t s) {
    if (!dst || on_device) return MPCB_OK;
    CK(cudaMemcpyAsync(dst, dev, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    return MPCB_OK;
  }
};

__global__ void gather_qp_out_kernel(int B, const int *__restrict__ status, const int *__restrict__ iter,
                                     int32_t *__restrict__ st_out, int32_t *__restrict__ it_out) {
  const int ln = blockIdx.x * blockDim.x + threadIdx.x;
  if (ln >= B) return;
  if (st_out) st_out[ln] = status[ln];
  if (it_out) it_out[ln] = iter[ln];
}

extern "C" int mpcb_qp_solve(mpcb_handle *h, int64_t B, const double *xhat, double *u0, int32_t *status, int32_t *iters,
                             int io_on_device) {
  if (!h || !xhat) return fail(MPCB_ERR_INVALID, "null argument");
  if (h->B == 0 || B != h->B) return fail(MPCB_ERR_STATE, "mpcb_batch_alloc(B) must precede mpcb_qp_solve with the same B");
  CK(cudaSetDevice(h->device));
  NvtxRange nvtx_("mpcb_qp_solve");
  Stage st(&h->stage_pool);
  const double *d_xhat;
  double *d_u0;
  int32_t *d_st, *d_it;
  RC(st.in(xhat, (size_t)6 * B, io_on_device, h->stream, &d_xhat));
  RC(st.out(u0, (size_t)2 * B, io_on_device, &d_u0));
  RC(st.out(status, (size_t)B, io_on_device, &d_st));
  RC(st.out(iters, (size_t)B, io_on_device, &d_it));
  AdmmArgs aa;
  PostArgs pa;
  fill_args(h, aa, pa, MODE_QP_ONLY);
  CK(cudaEventRecord(h->ev_t0, h->stream));
  const int pgrid = (int)((B + 127) / 128);
  if (h->generic_ok) {
    GenArgs g = h->gproto;
    g.mode = MODE_QP_ONLY;
    g.xhat = d_xhat;
    g.warm = 1;
    RC(launch_generic(h, g));
  } else if (!want_tile(h) && want_team(h)) {
    TeamArgs ta;
    fill_team_args(h, ta, MODE_QP_ONLY);
    ta.xhat = d_xhat;
    ta.warm = 1;
    RC(launch_team(h, ta));
  } else {
    CK(cudaMemsetAsync(h->cnt, 0, 12 * sizeof(int), h->stream));
    pa.cnt_cur = h->cnt + 4;
    pa.cnt_next = h->cnt;
    pa.list_next = h->list;
    qp_prepare_kernel<<<pgrid, 128, 0, h->stream>>>(pa, d_xhat);
    h->ctr.kernel_launches += 1;
    RC(run_rounds(h, aa, pa, 0, want_tile(h), !want_tile(h) && want_wave(h, true)));
  }
  if (d_u0) CK(cudaMemcpyAsync(d_u0, h->u0, (size_t)2 * B * 8, cudaMemcpyDeviceToDevice, h->stream));
  if (d_st || d_it) {
    gather_qp_out_kernel<<<pgrid, 128, 0, h->stream>>>((int)B, h->status, h->iter, d_st, d_it);
    h->ctr.kernel_launches += 1;
  }
  CK(cudaEventRecord(h->ev_t1, h->stream));
  RC(st.back(u0, d_u0, (size_t)2 * B, io_on_device, h->stream));
  RC(st.back(status, d_st, (size_t)B, io_on_device, h->stream));
  RC(st.back(iters, d_it, (size_t)B, io_on_device, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1));
  h->ctr.total_ms += ms;
  h->ctr.qp_solves += B;
  return pull_totals(h);
}

extern "C" int mpcb_qp_get_state(mpcb_handle *h, int64_t lane, double *x, double *z, double *y, double *rho) {
  if (!h || lane < 0 || lane >= h->B) return fail(MPCB_ERR_INVALID, "bad lane");
  CK(cudaSetDevice(h->device));
  const int n = h->hp.p.n, m = h->hp.p.m;
  CK(cudaStreamSynchronize(h->stream));
  if (x) CK(cudaMemcpy(x, h->xs + (size_t)lane * n, (size_t)n * 8, cudaMemcpyDeviceToHost));
  if (z) CK(cudaMemcpy(z, h->zs + (size_t)lane * m, (size_t)m * 8, cudaMemcpyDeviceToHost));
  if (y) CK(cudaMemcpy(y, h->ys + (size_t)lane * m, (size_t)m * 8, cudaMemcpyDeviceToHost));
  if (rho) CK(cudaMemcpy(rho, h->rho + lane, 8, cudaMemcpyDeviceToHost));
  return MPCB_OK;
}

// ------------------------------------------------------------------------------------------
extern "C" int mpcb_ukf_step(mpcb_handle *h, int64_t B, double *x, double *P, const double *u, const double *z,
                             int io_on_device) {
  if (!h || !x || !P || !u || !z || B < 1) return fail(MPCB_ERR_INVALID, "null argument");
  CK(cudaSetDevice(h->device));
  Stage st(&h->stage_pool);
  const double *dxi, *dPi, *du, *dz;
  RC(st.in((const double *)x, (size_t)6 * B, io_on_device, h->stream, &dxi));
  RC(st.in((const double *)P, (size_t)36 * B, io_on_device, h->stream, &dPi));
  RC(st.in(u, (size_t)2 * B, io_on_device, h->stream, &du));
  RC(st.in(z, (size_t)2 * B, io_on_device, h->stream, &dz));
  ukf_step_kernel<<<(int)((B + 63) / 64), 64, 0, h->stream>>>(h->sc, (int)B, (double *)dxi, (double *)dPi, du, dz);
  CK(cudaGetLastError());
  h->ctr.kernel_launches += 1;
  RC(st.back(x, dxi, (size_t)6 * B, io_on_device, h->stream));
  RC(st.back(P, dPi, (size_t)36 * B, io_on_device, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return MPCB_OK;
}

extern "C" int mpcb_plant_lin_step(mpcb_handle *h, int64_t B, double *x, const double *u, const double *w, int io_on_device) {
  if (!h || !x || !u || B < 1) return fail(MPCB_ERR_INVALID, "null argument");
  CK(cudaSetDevice(h->device));
  Stage st(&h->stage_pool);
  const double *dx, *du, *dw;
  RC(st.in((const double *)x, (size_t)4 * B, io_on_device, h->stream, &dx));
  RC(st.in(u, (size_t)2 * B, io_on_device, h->stream, &du));
  RC(st.in(w, (size_t)2 * B, io_on_device, h->stream, &dw));
  plant_lin_kernel<<<(int)((B + 127) / 128), 128, 0, h->stream>>>(h->sc, (int)B, (double *)dx, du, dw);
  CK(cudaGetLastError());
  h->ctr.kernel_launches += 1;
  RC(st.back(x, dx, (size_t)4 * B, io_on_device, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return MPCB_OK;
}

extern "C" int mpcb_plant_rk4(mpcb_handle *h, int64_t B, double *x, const double *u, const double *w, int nsub, double dt,
                              int io_on_device) {
  if (!h || !x || !u || B < 1 || nsub < 0) return fail(MPCB_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(h->device));
  Stage st(&h->stage_pool);
  const double *dx, *du, *dw;
  RC(st.in((const double *)x, (size_t)4 * B, io_on_device, h->stream, &dx));
  RC(st.in(u, (size_t)2 * B, io_on_device, h->stream, &du));
  RC(st.in(w, (size_t)2 * B, io_on_device, h->stream, &dw));
  plant_rk4_kernel<<<(int)((B + 127) / 128), 128, 0, h->stream>>>(h->sc, (int)B, (double *)dx, du, dw, nsub, dt);
  CK(cudaGetLastError());
  h->ctr.kernel_launches += 1;
  RC(st.back(x, dx, (size_t)4 * B, io_on_device, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return MPCB_OK;
}

// ------------------------------------------------------------------------------------------
static int simulate(mpcb_handle *h, int mode, int64_t B, int32_t nsteps, int32_t n_sub_total, int32_t ratio, double T_cont,
                    const double *x0, const double *noise, int32_t n_refresh, int32_t noise_hold_sub,
                    const mpcb_sim_out *out, int io_on_device) {
  if (!h || !x0) return fail(MPCB_ERR_INVALID, "null argument");
  if (h->B == 0 || B != h->B) return fail(MPCB_ERR_STATE, "mpcb_batch_alloc(B) must precede a simulation with the same B");
  if (nsteps < 0) return fail(MPCB_ERR_INVALID, "negative step count");
  if (h->sc.has_noise && (!noise || n_refresh < 1)) return fail(MPCB_ERR_INVALID, "has_noise problems need a noise tensor");
  if (h->sc.has_noise) {
    // the reference redraws every noise_length steps (trajectorySimulate.py:351-353; per noise_hold_sub substeps in
    // trajectorySimulateC.py:296-307): a tensor with fewer rows would silently reuse its last draw
    const int64_t need = (mode == MODE_CONTINUOUS)
                             ? (n_sub_total > 0 ? (int64_t)(n_sub_total - 1) / std::max(1, noise_hold_sub) + 1 : 1)
                             : (int64_t)nsteps / h->sc.noise_length + 1;
    if (n_refresh < need) {
      char b[160];
      snprintf(b, sizeof b, "noise tensor has %d rows, the run consumes %lld (one per hold interval)", n_refresh, (long long)need);
      return fail(MPCB_ERR_INVALID, b);
    }
  }
  NvtxRange nvtx_(mode == MODE_CONTINUOUS ? "mpcb_simulate_continuous" : "mpcb_simulate_discrete");
  CK(cudaSetDevice(h->device));
  RC(mpcb_batch_alloc(h, B));   // cold start: x = z = y = 0, rho = rho0 (a fresh osqp.setup, :242-245)
  static const mpcb_sim_out none = {};
  const mpcb_sim_out &o = out ? *out : none;
  const size_t T1 = (size_t)nsteps + 1;
  Stage st(&h->stage_pool);
  const double *d_x0, *d_noise;
  RC(st.in(x0, (size_t)4 * B, io_on_device, h->stream, &d_x0));
  RC(st.in(noise, noise ? (size_t)n_refresh * 2 * B : 0, io_on_device, h->stream, &d_noise));
  SimOutDev od;
  memset(&od, 0, sizeof od);
  od.T1 = (int)T1;
  RC(st.out(o.i_term, (size_t)B, io_on_device, &od.i_term));
  RC(st.out(o.is_success, (size_t)B, io_on_device, &od.is_success));
  RC(st.out(o.final_dist, (size_t)B, io_on_device, &od.final_dist));
  RC(st.out(o.x_true, 4 * T1 * B, io_on_device, &od.x_true));
  RC(st.out(o.x_est, 6 * T1 * B, io_on_device, &od.x_est));
  RC(st.out(o.ctrl, 2 * T1 * B, io_on_device, &od.ctrl));
  RC(st.out(o.ctrlr_seq, (T1 - 1) * B, io_on_device, &od.ctrlr_seq));
  RC(st.out(o.status, (T1 - 1) * B, io_on_device, &od.status));
  RC(st.out(o.iters, (T1 - 1) * B, io_on_device, &od.iters));
  RC(st.out(o.u_raw, 2 * (T1 - 1) * B, io_on_device, &od.u_raw));
  RC(st.out(o.ukf_clamped, (size_t)B, io_on_device, &od.ukf_clamped));
  RC(st.out(o.rho, (T1 - 1) * B, io_on_device, &od.rho_hist));
  od.fd_all = h->lane_fd;
  const size_t NS = (mode == MODE_CONTINUOUS) ? (size_t)n_sub_total : 0;
  od.NS = (int)NS;
  if (mode == MODE_CONTINUOUS) {
    RC(st.out(o.x_true_sub, 4 * NS * B, io_on_device, &od.x_true_sub));
    RC(st.out(o.ctrl_sub, 2 * NS * B, io_on_device, &od.ctrl_sub));
    RC(st.out(o.ctrlr_sub, NS * B, io_on_device, &od.ctrlr_sub));
    if (od.x_true_sub) CK(cudaMemsetAsync(od.x_true_sub, 0xff, 4 * NS * B * 8, h->stream));
    if (od.ctrl_sub) CK(cudaMemsetAsync(od.ctrl_sub, 0xff, 2 * NS * B * 8, h->stream));
    if (od.ctrlr_sub) CK(cudaMemsetAsync(od.ctrlr_sub, 0, NS * B, h->stream));
  }
  // lanes that stop early leave the rest of their telemetry columns undefined in the reference
  // (np.empty, :258-262); here they read NaN / 0
  if (od.x_true) CK(cudaMemsetAsync(od.x_true, 0xff, 4 * T1 * B * 8, h->stream));
  if (od.x_est) CK(cudaMemsetAsync(od.x_est, 0xff, 6 * T1 * B * 8, h->stream));
  if (od.ctrl) CK(cudaMemsetAsync(od.ctrl, 0xff, 2 * T1 * B * 8, h->stream));
  if (od.u_raw) CK(cudaMemsetAsync(od.u_raw, 0xff, 2 * (T1 - 1) * B * 8, h->stream));
  if (od.ctrlr_seq) CK(cudaMemsetAsync(od.ctrlr_seq, 0, (T1 - 1) * B, h->stream));
  if (od.status) CK(cudaMemsetAsync(od.status, 0, (T1 - 1) * B, h->stream));
  if (od.iters) CK(cudaMemsetAsync(od.iters, 0, (T1 - 1) * B * 2, h->stream));
  if (od.rho_hist) CK(cudaMemsetAsync(od.rho_hist, 0xff, (T1 - 1) * B * 8, h->stream));

  AdmmArgs aa;
  PostArgs pa;
  fill_args(h, aa, pa, mode);
  pa.out = od;
  pa.nsteps = nsteps;
  pa.ratio = ratio;
  pa.n_sub_total = n_sub_total;
  pa.noise_hold_sub = std::max(1, noise_hold_sub);
  pa.T_cont = T_cont;
  pa.noise_in = d_noise;
  pa.n_refresh = n_refresh;
  CK(cudaEventRecord(h->ev_t0, h->stream));
  CK(cudaMemsetAsync(h->d_stats, 0, MPCB_NSTATS * sizeof(double), h->stream));
  if (h->generic_ok) {
    GenArgs g = h->gproto;
    g.mode = mode;
    g.nsteps = nsteps;
    g.ratio = ratio;
    g.n_sub_total = n_sub_total;
    g.noise_hold_sub = std::max(1, noise_hold_sub);
    g.T_cont = T_cont;
    g.out = od;
    g.x0 = d_x0;
    g.noise_in = d_noise;
    g.n_refresh = n_refresh;
    g.warm = 0;
    RC(launch_generic(h, g));
  } else if (!want_tile(h) && !want_wave(h, true) && want_team(h) && mode == MODE_DISCRETE) {
    TeamArgs ta;
    fill_team_args(h, ta, MODE_DISCRETE);
    ta.nsteps = nsteps;
    ta.out = od;
    ta.x0 = d_x0;
    ta.noise_in = d_noise;
    ta.n_refresh = n_refresh;
    ta.warm = 0;
    RC(launch_team(h, ta));
  } else {
    pa.cnt_cur = h->cnt + 4;
    pa.cnt_next = h->cnt;
    pa.list_next = h->list;
    const bool can_resume = mode == MODE_DISCRETE && h->team_ok && !getenv("MPCB_NO_RESUME");
    {
      // rounds on a block that only counts re-typed rows (wave, tile, block kernels): lanes whose solve would have them go to
      // the team kernel instead (PostArgs::defer_below).  |p^ - r|_1 E_r < RHO_TOL for some velocity-bound row r.
      const bool rounds_on_team = !want_tile(h) && !(want_wave(h, true) && mode == MODE_DISCRETE) && want_team(h);
      pa.defer_below = 0.0;
      pa.cnt_def = h->cnt + 8;
      pa.list_def = h->list + (size_t)8 * B;
      if (can_resume && !rounds_on_team && h->team_tm && h->thdr.r3ok && !getenv("MPCB_NO_DEFER")) {
        const mpcb_problem &pp = h->hp.p;
        double emin = 1e300;              // the row with the smallest scale factor is re-typed first
        for (int k = 0; k <= pp.Nb; ++k) emin = std::min(emin, h->hp.E[4 * (pp.Nx + 1) + 5 * k + 3]);
        if (emin > 0.0 && emin < 1e300) pa.defer_below = MPCB_RHO_TOL / emin * (1.0 + 1e-12);
      }
    }
    const int pgrid = (int)((B + 127) / 128);
    init_kernel<<<pgrid, 128, 0, h->stream>>>(pa, d_x0);
    CK(cudaGetLastError());
    h->ctr.kernel_launches += 1;
    TeamArgs tr;
    if (can_resume) {
      fill_team_args(h, tr, MODE_RESUME);
      tr.nsteps = nsteps;
      tr.out = od;
      tr.noise_in = d_noise;
      tr.n_refresh = n_refresh;
      tr.warm = 1;
      tr.ls = h->ls;
      tr.par = h->par;
      tr.lane_state = h->lane_state;
      tr.visit_iters = 0;
    }
    long below = 10L * h->num_sms * h->team_ctas;
    if (const char *e = getenv("MPCB_RESUME_BELOW")) below = atol(e);
    // Solver block of the rounds.  Discrete simulator, >= wave_min_lanes: multi-RHS wave kernel, 25 iterations per round, the last
    // lanes handed to the team kernel.  Continuous simulator: the team kernel in list mode (every round's solves run to
    // completion, one round per control step) -- measured on config 3 (65 536 lanes): 2.43 s per step against 3.13 s with wave
    // rounds, which pay a post_kernel pass (500 RK4 substeps per lane) and a host round trip per 25 iterations.
    const bool forced_wave = getenv("MPCB_SOLVER") && strcmp(getenv("MPCB_SOLVER"), "wave") == 0;
    const bool use_wave = !want_tile(h) && want_wave(h, true) && (mode == MODE_DISCRETE || forced_wave);
    h->wave_warps_now = ((B + 31) / 32 <= h->num_sms && !getenv("MPCB_NO_WAVE4")) ? 4 : WAVE_WARPS;
    RC(run_rounds(h, aa, pa, 0, want_tile(h), use_wave, can_resume ? &tr : nullptr, below));
    if (pa.defer_below > 0.0) {          // lanes that left the rounds because OSQP re-types their rows: team kernel, to the end
      CK(cudaMemcpyAsync(h->h_cnt, h->cnt + 8, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      long ndef = 0;
      for (int v = 0; v < 4; ++v) ndef += h->h_cnt[v];
      if (ndef > 0) {
        TeamArgs ta = tr;
        ta.cnt = h->cnt + 8;
        ta.list = h->list + (size_t)8 * B;
        CK(cudaMemsetAsync(h->d_queue, 0, sizeof(int), h->stream));
        const int tgrid = (int)std::min<long>(ndef, (long)h->num_sms * h->team_ctas);
        ((team_fn)h->team_fn_ptr)<<<tgrid, h->team_threads, h->team_smem, h->stream>>>(ta);
        CK(cudaGetLastError());
        h->ctr.kernel_launches += 1;
        h->ctr.admm_launches += 1;
      }
    }
    finalize_kernel<<<pgrid, 128, 0, h->stream>>>(pa, h->d_stats, h->flip);
    CK(cudaGetLastError());
    h->ctr.kernel_launches += 1;
  }
  stats_fd_kernel<<<1, 1024, 0, h->stream>>>(h->lane_fd, (int)B, h->d_stats);
  CK(cudaGetLastError());
  h->ctr.kernel_launches += 1;
  CK(cudaEventRecord(h->ev_t1, h->stream));
  RC(st.back(o.i_term, od.i_term, (size_t)B, io_on_device, h->stream));
  RC(st.back(o.is_success, od.is_success, (size_t)B, io_on_device, h->stream));
  RC(st.back(o.final_dist, od.final_dist, (size_t)B, io_on_device, h->stream));
  RC(st.back(o.x_true, od.x_true, 4 * T1 * B, io_on_device, h->stream));
  RC(st.back(o.x_est, od.x_est, 6 * T1 * B, io_on_device, h->stream));
  RC(st.back(o.ctrl, od.ctrl, 2 * T1 * B, io_on_device, h->stream));
  RC(st.back(o.ctrlr_seq, od.ctrlr_seq, (T1 - 1) * B, io_on_device, h->stream));
  RC(st.back(o.status, od.status, (T1 - 1) * B, io_on_device, h->stream));
  RC(st.back(o.iters, od.iters, (T1 - 1) * B, io_on_device, h->stream));
  RC(st.back(o.u_raw, od.u_raw, 2 * (T1 - 1) * B, io_on_device, h->stream));
  RC(st.back(o.ukf_clamped, od.ukf_clamped, (size_t)B, io_on_device, h->stream));
  RC(st.back(o.rho, od.rho_hist, (T1 - 1) * B, io_on_device, h->stream));
  if (mode == MODE_CONTINUOUS) {
    RC(st.back(o.x_true_sub, od.x_true_sub, 4 * NS * B, io_on_device, h->stream));
    RC(st.back(o.ctrl_sub, od.ctrl_sub, 2 * NS * B, io_on_device, h->stream));
    RC(st.back(o.ctrlr_sub, od.ctrlr_sub, NS * B, io_on_device, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1));
  h->ctr.total_ms += ms;
  double stats[MPCB_NSTATS];
  CK(cudaMemcpy(stats, h->d_stats, sizeof stats, cudaMemcpyDeviceToHost));
  h->ctr.qp_solves += (int64_t)stats[5];
  h->ctr.flip_lanes += (int64_t)stats[7];
  h->sim_done = true;
  const int64_t before = h->ctr.admm_iterations;
  RC(pull_totals(h));
  h->last_sim_iterations = h->ctr.admm_iterations - before;
  return MPCB_OK;
}

extern "C" int mpcb_simulate_discrete(mpcb_handle *h, int64_t B, int32_t nsteps, const double *x0, const double *noise,
                                      int32_t n_refresh, const mpcb_sim_out *out, int io_on_device) {
  return simulate(h, MODE_DISCRETE, B, nsteps, 0, 1, 0.0, x0, noise, n_refresh, 1, out, io_on_device);
}

extern "C" int mpcb_simulate_continuous(mpcb_handle *h, int64_t B, int32_t n_sub_total, int32_t ratio, double T_cont,
                                        const double *x0, const double *noise, int32_t n_refresh, int32_t noise_hold_sub,
                                        const mpcb_sim_out *out, int io_on_device) {
  if (ratio < 1 || n_sub_total < 0 || !(T_cont > 0.0)) return fail(MPCB_ERR_INVALID, "bad continuous-time grid");
  const int32_t n_samples = n_sub_total / ratio;   // nsimD = int(T_final / T) (trajectorySimulateC.py:55-58)
  return simulate(h, MODE_CONTINUOUS, B, n_samples, n_sub_total, ratio, T_cont, x0, noise, n_refresh, noise_hold_sub, out,
                  io_on_device);
}

extern "C" int mpcb_noise_fill(mpcb_handle *h, int64_t B, int32_t n_refresh, double sigma_x, double sigma_y, uint64_t seed,
                               uint64_t lane_offset, double *noise, uint32_t *raw, int io_on_device) {
  if (!h || !noise || B < 1 || n_refresh < 1 || n_refresh > 65535) return fail(MPCB_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(h->device));
  Stage st(&h->stage_pool);
  double *d_noise;
  uint32_t *d_raw = nullptr;
  RC(st.out(noise, (size_t)n_refresh * 2 * B, io_on_device, &d_noise));
  if (raw) RC(st.out(raw, (size_t)n_refresh * 4 * B, io_on_device, &d_raw));
  const dim3 grid((unsigned)((B + 127) / 128), (unsigned)n_refresh);
  noise_fill_kernel<<<grid, 128, 0, h->stream>>>((int)B, n_refresh, sigma_x, sigma_y, seed, lane_offset, d_noise, d_raw);
  CK(cudaGetLastError());
  h->ctr.kernel_launches += 1;
  RC(st.back(noise, d_noise, (size_t)n_refresh * 2 * B, io_on_device, h->stream));
  if (raw) RC(st.back(raw, d_raw, (size_t)n_refresh * 4 * B, io_on_device, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return MPCB_OK;
}

// ncclAllReduce resolved at run time: libmpcb.so has no link-time NCCL dependency
typedef int (*nccl_allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
static nccl_allreduce_fn resolve_nccl_allreduce() {
  static nccl_allreduce_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *sym = dlsym(RTLD_DEFAULT, "ncclAllReduce");        // already loaded (e.g. by torch): use that copy
    if (!sym) {
      void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
      if (lib) sym = dlsym(lib, "ncclAllReduce");
    }
    fn = (nccl_allreduce_fn)sym;
  }
  return fn;
}

extern "C" int mpcb_allreduce_stats(mpcb_handle *h, void *nccl_comm, double *stats_out) {
  if (!h || !nccl_comm || !stats_out) return fail(MPCB_ERR_INVALID, "null argument");
  if (!h->sim_done) return fail(MPCB_ERR_STATE, "mpcb_allreduce_stats needs a finished simulation");
  nccl_allreduce_fn ar = resolve_nccl_allreduce();
  if (!ar) return fail(MPCB_ERR_STATE, "libnccl.so.2 not found (ncclAllReduce could not be resolved)");
  CK(cudaSetDevice(h->device));
  NvtxRange nvtx_("mpcb_allreduce_stats");
  const double iters = (double)h->last_sim_iterations;
  CK(cudaMemcpyAsync(h->d_stats + 6, &iters, 8, cudaMemcpyHostToDevice, h->stream));
  const int rc = ar(h->d_stats, h->d_stats + MPCB_NSTATS, MPCB_NSTATS, /*ncclDouble*/ 8, /*ncclSum*/ 0, nccl_comm, h->stream);
  if (rc != 0) {
    char b[96];
    snprintf(b, sizeof b, "ncclAllReduce failed with ncclResult_t %d", rc);
    return fail(MPCB_ERR_CUDA, b);
  }
  CK(cudaMemcpyAsync(stats_out, h->d_stats + MPCB_NSTATS, MPCB_NSTATS * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return MPCB_OK;
}

extern "C" int mpcb_stats(mpcb_handle *h, int64_t B, double *stats_out, int io_on_device) {
  if (!h || !stats_out) return fail(MPCB_ERR_INVALID, "null argument");
  if (!h->sim_done || B != h->B) return fail(MPCB_ERR_STATE, "mpcb_stats needs a finished simulation of the same batch");
  CK(cudaSetDevice(h->device));
  double s[MPCB_NSTATS];
  CK(cudaMemcpy(s, h->d_stats, sizeof s, cudaMemcpyDeviceToHost));
  s[6] = (double)h->last_sim_iterations;
  if (io_on_device) CK(cudaMemcpy(stats_out, s, sizeof s, cudaMemcpyHostToDevice));
  else memcpy(stats_out, s, sizeof s);
  return MPCB_OK;
}
