"""Python face of the C-ABI engine: one :class:`Engine` = one ``mpcb_handle`` = one problem family
on one GPU.  Arrays are float64 structure-of-arrays ``[field, B]`` (lane index fastest), either
numpy (host pointers; the library stages them itself) or torch CUDA tensors (device pointers,
zero copies).  PyTorch is used only to own device memory.

Reference seams: the OSQP protocol ``setup / update / solve`` (``src/trajectorySimulate.py:242-245,
296,340-348``), ``kf.predict / kf.update`` (``:333-335``), the plant steps (``:324``;
``src/trajectorySimulateC.py:372-380``) and the two closed loops (``:285-356``; ``…SimulateC.py:325-410``).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .mpcsim import BatchSimRun
from .problem import Problem

_RECORD_ALL = ("x_true", "x_est", "ctrl", "ctrlr_seq", "status", "iters", "u_raw", "rho")


def _is_torch(a) -> bool:
    return type(a).__module__.startswith("torch")


class _Arrays:
    """Allocates outputs of the same kind (numpy / torch-cuda) as the inputs and hands out pointers."""

    def __init__(self, like, device_index: int, host_cache=None):
        self.on_device = _is_torch(like) and like.is_cuda
        self.dev = device_index
        self.host_cache = host_cache        # dict of pinned host arrays reused across calls, or None
        if _is_torch(like) and not like.is_cuda:
            raise TypeError("torch inputs must be CUDA tensors (use numpy for host arrays)")
        self._keep = []

    def ptr(self, a, dtype=np.float64, shape=None):
        if a is None:
            return None
        if self.on_device:
            import torch
            if not (_is_torch(a) and a.is_cuda):
                raise TypeError("mixing host and device arrays in one call")
            want = {np.float64: torch.float64, np.int32: torch.int32}[dtype]
            if a.dtype != want or not a.is_contiguous():
                raise TypeError(f"device arrays must be contiguous {want}")
            if a.device.index != self.dev:
                raise ValueError(f"tensor on cuda:{a.device.index}, engine on cuda:{self.dev}")
            if shape is not None and tuple(a.shape) != tuple(shape):
                raise ValueError(f"expected shape {tuple(shape)}, got {tuple(a.shape)}")
            self._keep.append(a)
            return C.c_void_p(a.data_ptr())
        a = np.ascontiguousarray(a, dtype=dtype)
        if shape is not None and a.shape != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {a.shape}")
        self._keep.append(a)
        return C.c_void_p(a.ctypes.data)

    def new(self, shape, dtype=np.float64, name=None):
        if not self.on_device and self.host_cache is not None and name is not None:
            key = (name, tuple(shape), np.dtype(dtype).str)
            a = self.host_cache.get(key)
            if a is None:
                import torch
                tdt = {np.float64: torch.float64, np.int32: torch.int32, np.uint8: torch.uint8, np.int8: torch.int8,
                       np.int16: torch.int16}[dtype]
                a = torch.empty(tuple(shape), dtype=tdt, pin_memory=True).numpy()
                self.host_cache[key] = a
            return a, C.c_void_p(a.ctypes.data)
        if self.on_device:
            import torch
            tdt = {np.float64: torch.float64, np.int32: torch.int32, np.uint8: torch.uint8, np.int8: torch.int8,
                   np.int16: torch.int16}[dtype]
            t = torch.empty(tuple(shape), dtype=tdt, device=f"cuda:{self.dev}")
            return t, C.c_void_p(t.data_ptr())
        a = np.empty(tuple(shape), dtype=dtype)
        return a, C.c_void_p(a.ctypes.data)


def fill_problem_struct(problem: Problem):
    """``struct mpcb_problem`` (include/mpcb.h) for a :class:`Problem`.  Returns the ctypes struct and the list of numpy
    arrays its pointers refer to (keep it alive as long as the struct is used)."""
    p, st = problem, problem.settings
    cp = _lib.MpcbProblem()
    cp.Nx, cp.Nc, cp.Nb, cp.n, cp.m = p.Nx, p.Nc, p.Nb, p.n, p.m
    cp.in_track, cp.delta_v, cp.is_reject, cp.has_noise = int(p.in_track), int(p.delta_v), int(p.is_reject), int(p.has_noise)
    cp.noise_length = int(p.noise_length)
    cp.estimator = int(p.estimator)
    cp.rho0, cp.sigma, cp.alpha = st.rho, st.sigma, st.alpha
    cp.eps_abs, cp.eps_rel, cp.eps_prim_inf = st.eps_abs, st.eps_rel, st.eps_prim_inf
    cp.adaptive_rho_tolerance = st.adaptive_rho_tolerance
    cp.max_iter, cp.check_termination = st.max_iter, st.check_termination
    cp.adaptive_rho, cp.adaptive_rho_interval = int(st.adaptive_rho), st.adaptive_rho_interval
    tables = []

    def fixed(name, arr, cnt):
        a = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
        assert a.size == cnt, (name, a.size, cnt)
        getattr(cp, name)[:] = a.tolist()

    fixed("Ad", p.Ad, 16); fixed("Bd", p.Bd, 8); fixed("Ao", p.Ao, 36); fixed("Bou", p.Bou, 12)
    fixed("Qw", p.Qw, 36); fixed("Kpf", p.Kpf, 8); fixed("Kif", p.Kif, 2); fixed("xr", p.xr, 4)
    cp.umax0, cp.r_p, cp.r_tol = p.umax0, p.r_p, p.r_tol
    cp.suc_dist, cp.suc_ang_deg, cp.mean_mtn, cp.T = p.suc_dist, p.suc_ang, p.mean_mtn, p.T

    def table(name, arr, dtype=np.float64):
        a = np.ascontiguousarray(arr, dtype=dtype)
        tables.append(a)
        ptr_t = _lib.c_double_p if dtype == np.float64 else _lib.c_int32_p
        setattr(cp, name, a.ctypes.data_as(ptr_t))

    cp.has_debris, cp.scaling = int(p.has_debris), int(st.scaling)
    if p.has_debris:
        table("P_u", p.P); table("q_u", p.q); table("A_u", p.A); table("l_u", p.l); table("u_u", p.u)
        cp.debris_center[:] = np.asarray(p.debris_center, float).tolist()
        cp.debris_side, cp.debris_detect = float(p.debris_side), float(p.debris_detect)
        cp.debris_verts[:] = np.asarray(p.debris_verts, float).reshape(-1).tolist()
    else:
        table("P_s", p.P_s); table("q_s", p.q_s); table("A_s", p.A_s); table("l_s", p.l_s); table("u_s", p.u_s)
        table("D", p.D); table("E", p.E); table("ctype", p.ctype, np.int32); table("V", p.V); table("lam", p.lam)
        # unscaled data as well: the multi-RHS (wave) solver block iterates in unscaled variables
        table("P_u", p.P); table("q_u", p.q); table("A_u", p.A); table("l_u", p.l); table("u_u", p.u)
    cp.K_dead[:] = np.asarray(p.K_dead, float).reshape(-1).tolist()
    cp.Ki_dead[:] = np.asarray(p.Ki_dead, float).reshape(-1).tolist()
    cp.c = p.c
    return cp, tables


class Engine:
    """``pin_outputs=True`` returns host results in page-locked arrays that the engine keeps and
    OVERWRITES on the next call with the same shapes (fast D2H for repeated batches; copy what you keep)."""

    def __init__(self, problem: Problem, device: int = 0, pin_outputs: bool = False):
        self._host_cache = {} if pin_outputs else None
        self.lib = _lib.load()
        self.problem = problem
        self.device = int(device)
        self._h = C.c_void_p()
        self.B = 0
        cp, self._tables = fill_problem_struct(problem)
        _lib.check(self.lib.mpcb_create(C.byref(cp), self.device, C.byref(self._h)))

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if self._h:
            self.lib.mpcb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def batch_alloc(self, B: int):
        """Allocate per-lane state and cold-start every lane's solver (a fresh ``osqp.setup``)."""
        _lib.check(self.lib.mpcb_batch_alloc(self._h, int(B)))
        self.B = int(B)

    def set_timing(self, on: bool):
        _lib.check(self.lib.mpcb_set_timing(self._h, int(on)))

    def counters(self) -> dict:
        c = _lib.MpcbCounters()
        _lib.check(self.lib.mpcb_get_counters(self._h, C.byref(c)))
        return {k: getattr(c, k) for k, _ in c._fields_}

    def solver_blocks(self) -> dict:
        """Which solver blocks this problem family can run on (``mpcb_solver_blocks``)."""
        m = int(self.lib.mpcb_solver_blocks(self._h))
        return {"block": bool(m & 1), "team": bool(m & 2), "tile": bool(m & 4), "wave": bool(m & 8), "generic": bool(m & 16)}

    @property
    def stream(self) -> int:
        return int(self.lib.mpcb_stream(self._h) or 0)

    # ------------------------------------------------------------------ unit seams
    def qp_solve(self, xhat):
        """``prob.update(...)`` x2 + ``prob.solve()`` for every lane; warm-started from the lane's
        previous solve.  ``xhat[6, B]`` -> ``(u0[2, B], status[B], iters[B])``."""
        B = xhat.shape[1]
        if B != self.B:
            self.batch_alloc(B)
        ar = _Arrays(xhat, self.device)
        self._order_after_caller(ar)
        u0, pu = ar.new((2, B))
        st, ps = ar.new((B,), np.int32)
        it, pi = ar.new((B,), np.int32)
        _lib.check(self.lib.mpcb_qp_solve(self._h, B, ar.ptr(xhat, shape=(6, B)), pu, ps, pi, int(ar.on_device)))
        return u0, st, it

    def qp_state(self, lane: int):
        p = self.problem
        x, z, y, rho = np.empty(p.n), np.empty(p.m), np.empty(p.m), np.empty(1)
        _lib.check(self.lib.mpcb_qp_get_state(self._h, int(lane), x.ctypes.data, z.ctypes.data, y.ctypes.data, rho.ctypes.data))
        return x, z, y, float(rho[0])

    def ukf_step(self, x, P, u, z):
        """``kf.predict(u); kf.update(z)`` on copies of ``x[6, B]``, ``P[36, B]``; returns the new (x, P)."""
        B = x.shape[1]
        ar = _Arrays(x, self.device)
        self._order_after_caller(ar)
        x = x.clone() if ar.on_device else np.array(x, dtype=np.float64, order="C")
        P = P.clone() if ar.on_device else np.array(P, dtype=np.float64, order="C")
        _lib.check(self.lib.mpcb_ukf_step(self._h, B, ar.ptr(x, shape=(6, B)), ar.ptr(P, shape=(36, B)),
                                          ar.ptr(u, shape=(2, B)), ar.ptr(z, shape=(2, B)), int(ar.on_device)))
        return x, P

    def noise_fill(self, B: int, n_refresh: int, seed: int, lane_offset: int = 0, on_device: bool = False, raw: bool = False):
        """Disturbance tensor ``[n_refresh, 2, B]`` drawn on the GPU (Philox4x32-10, ``mpcb_noise_fill``): ``sigma`` of the
        problem times N(0, 1), one generator block per (lane + lane_offset, refresh).  ``on_device``: return a torch CUDA
        tensor (stays in HBM for the simulators) instead of a numpy array.  ``raw``: also return the generator words."""
        p = self.problem
        if on_device:
            import torch
            dev = torch.device("cuda", self.device)
            noise = torch.empty((n_refresh, 2, B), dtype=torch.float64, device=dev)
            words = torch.empty((n_refresh, 4, B), dtype=torch.int32, device=dev) if raw else None
            np_, wp_ = noise.data_ptr(), (words.data_ptr() if raw else None)
        else:
            noise = np.empty((n_refresh, 2, B))
            words = np.empty((n_refresh, 4, B), dtype=np.uint32) if raw else None
            np_, wp_ = noise.ctypes.data, (words.ctypes.data if raw else None)
        _lib.check(self.lib.mpcb_noise_fill(self._h, B, int(n_refresh), float(p.sig[0]), float(p.sig[1]), int(seed) & (2 ** 64 - 1),
                                            int(lane_offset), np_, wp_, int(on_device)))
        return (noise, words) if raw else noise

    def plant_lin_step(self, x, u, w=None):
        B = x.shape[1]
        ar = _Arrays(x, self.device)
        self._order_after_caller(ar)
        x = x.clone() if ar.on_device else np.array(x, dtype=np.float64, order="C")
        _lib.check(self.lib.mpcb_plant_lin_step(self._h, B, ar.ptr(x, shape=(4, B)), ar.ptr(u, shape=(2, B)),
                                                ar.ptr(w, shape=(2, B)) if w is not None else None, int(ar.on_device)))
        return x

    def plant_rk4(self, x, u, w, nsub: int, dt: float):
        B = x.shape[1]
        ar = _Arrays(x, self.device)
        self._order_after_caller(ar)
        x = x.clone() if ar.on_device else np.array(x, dtype=np.float64, order="C")
        _lib.check(self.lib.mpcb_plant_rk4(self._h, B, ar.ptr(x, shape=(4, B)), ar.ptr(u, shape=(2, B)),
                                           ar.ptr(w, shape=(2, B)) if w is not None else None, int(nsub), float(dt),
                                           int(ar.on_device)))
        return x

    # ------------------------------------------------------------------ closed loops
    def _sim_outputs(self, ar: _Arrays, B: int, T1: int, record: Sequence[str], NS: int = 0):
        out = _lib.MpcbSimOut()
        res = {}
        spec = {
            "i_term": ((B,), np.int32), "is_success": ((B,), np.int32), "final_dist": ((B,), np.float64),
            "ukf_clamped": ((B,), np.int32),
            "x_true": ((4, T1, B), np.float64), "x_est": ((6, T1, B), np.float64), "ctrl": ((2, T1, B), np.float64),
            "ctrlr_seq": ((T1 - 1, B), np.uint8), "status": ((T1 - 1, B), np.int8), "iters": ((T1 - 1, B), np.int16),
            "u_raw": ((2, T1 - 1, B), np.float64), "rho": ((T1 - 1, B), np.float64),
            "x_true_sub": ((4, NS, B), np.float64), "ctrl_sub": ((2, NS, B), np.float64), "ctrlr_sub": ((NS, B), np.uint8),
        }
        for name in ("i_term", "is_success", "final_dist", "ukf_clamped") + tuple(record):
            shape, dt = spec[name]
            arr, p = ar.new(shape, dt, name)
            res[name] = arr
            setattr(out, name, p)
        return out, res

    def _wrap(self, res: dict, B: int) -> BatchSimRun:
        stats = np.empty(_lib.NSTATS)
        _lib.check(self.lib.mpcb_stats(self._h, B, stats.ctypes.data, 0))
        keys = ("sum_final_dist", "sum_final_dist_sq", "n_success", "n_lanes", "sum_i_term", "qp_solves", "admm_iterations",
                "flip_lanes", "ukf_clamped_lanes", "early_term_lanes")
        return BatchSimRun(i_term=res["i_term"], isSuccess=res["is_success"], final_dist=res["final_dist"],
                           x_true=res.get("x_true"), x_est=res.get("x_est"), ctrl_hist=res.get("ctrl"),
                           ctrlr_seq=res.get("ctrlr_seq"), status=res.get("status"), iters=res.get("iters"),
                           u_raw=res.get("u_raw"), rho=res.get("rho"), ukf_clamped=res["ukf_clamped"], x_true_sub=res.get("x_true_sub"),
                           ctrl_sub=res.get("ctrl_sub"), ctrlr_sub=res.get("ctrlr_sub"), stats=dict(zip(keys, stats.tolist())),
                           stats_vec=stats)

    def _order_after_caller(self, ar: "_Arrays"):
        """Device-pointer calls: order the engine's stream after torch's current stream, so inputs still being produced
        by torch kernels are complete before the engine reads them (``mpcb_wait_stream``; stream contract in mpcb.h).
        Outputs need no such step: every call returns after the engine's stream has drained."""
        if ar.on_device:
            import torch
            _lib.check(self.lib.mpcb_wait_stream(self._h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def simulate_discrete(self, x0, noise, nsteps: int, record: Sequence[str] = _RECORD_ALL) -> BatchSimRun:
        """``trajectorySimulate`` for B lanes.  ``x0[4, B]``; ``noise[R, 2, B]`` sigma-scaled position
        disturbances (row r is applied from step ``r*noise_length``; ``None`` for noise-free problems); ``nsteps``
        control steps; ``record`` selects telemetry."""
        B = x0.shape[1]
        if B != self.B:
            self.batch_alloc(B)
        ar = _Arrays(x0, self.device, self._host_cache)
        self._order_after_caller(ar)
        out, res = self._sim_outputs(ar, B, nsteps + 1, record)
        R = 0 if noise is None else noise.shape[0]
        _lib.check(self.lib.mpcb_simulate_discrete(self._h, B, int(nsteps), ar.ptr(x0, shape=(4, B)),
                                                   ar.ptr(noise, shape=(R, 2, B)) if noise is not None else None, R,
                                                   C.byref(out), int(ar.on_device)))
        return self._wrap(res, B)

    def simulate_continuous(self, x0, noise, n_sub_total: int, ratio: int, T_cont: float, noise_hold_sub: int,
                            record: Sequence[str] = _RECORD_ALL) -> BatchSimRun:
        """``trajectorySimulateC`` for B lanes (RK4 at ``h = T_cont``); telemetry at the sample instants."""
        B = x0.shape[1]
        if B != self.B:
            self.batch_alloc(B)
        ar = _Arrays(x0, self.device, self._host_cache)
        self._order_after_caller(ar)
        T1 = n_sub_total // ratio + 1
        out, res = self._sim_outputs(ar, B, T1, record, int(n_sub_total))
        R = 0 if noise is None else noise.shape[0]
        _lib.check(self.lib.mpcb_simulate_continuous(self._h, B, int(n_sub_total), int(ratio), float(T_cont),
                                                     ar.ptr(x0, shape=(4, B)),
                                                     ar.ptr(noise, shape=(R, 2, B)) if noise is not None else None, R,
                                                     int(noise_hold_sub), C.byref(out), int(ar.on_device)))
        return self._wrap(res, B)
