"""ctypes binding of ``libmpcb.so`` (C ABI in ``include/mpcb.h``).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a) into
``mpc_arpo_project_b200/lib/libmpcb.so``.  There is no CPU fallback: if the library is
missing or no CUDA device is present the engine raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MPCB_LIB: load another build of the same ABI (A/B timing of kernel variants on one box); default is the in-tree library
LIB_PATH = os.environ.get("MPCB_LIB") or os.path.join(_HERE, "lib", "libmpcb.so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class MpcbProblem(C.Structure):
    """``struct mpcb_problem`` (include/mpcb.h)."""
    _fields_ = [
        ("Nx", C.c_int32), ("Nc", C.c_int32), ("Nb", C.c_int32),
        ("n", C.c_int32), ("m", C.c_int32),
        ("in_track", C.c_int32), ("delta_v", C.c_int32), ("is_reject", C.c_int32), ("has_noise", C.c_int32),
        ("noise_length", C.c_int32), ("estimator", C.c_int32),
        ("rho0", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double), ("eps_abs", C.c_double),
        ("eps_rel", C.c_double), ("eps_prim_inf", C.c_double), ("adaptive_rho_tolerance", C.c_double),
        ("max_iter", C.c_int32), ("check_termination", C.c_int32), ("adaptive_rho", C.c_int32),
        ("adaptive_rho_interval", C.c_int32),
        ("Ad", C.c_double * 16), ("Bd", C.c_double * 8), ("Ao", C.c_double * 36), ("Bou", C.c_double * 12),
        ("Qw", C.c_double * 36), ("Kpf", C.c_double * 8), ("Kif", C.c_double * 2), ("xr", C.c_double * 4),
        ("umax0", C.c_double), ("r_p", C.c_double), ("r_tol", C.c_double), ("suc_dist", C.c_double),
        ("suc_ang_deg", C.c_double), ("mean_mtn", C.c_double), ("T", C.c_double),
        ("P_s", c_double_p), ("q_s", c_double_p), ("A_s", c_double_p), ("l_s", c_double_p), ("u_s", c_double_p),
        ("D", c_double_p), ("E", c_double_p), ("c", C.c_double),
        ("ctype", c_int32_p), ("V", c_double_p), ("lam", c_double_p),
        ("has_debris", C.c_int32), ("scaling", C.c_int32),
        ("debris_center", C.c_double * 2), ("debris_side", C.c_double), ("debris_detect", C.c_double),
        ("debris_verts", C.c_double * 8), ("K_dead", C.c_double * 8), ("Ki_dead", C.c_double * 2),
        ("P_u", c_double_p), ("q_u", c_double_p), ("A_u", c_double_p), ("l_u", c_double_p), ("u_u", c_double_p),
    ]


class MpcbSimOut(C.Structure):
    """``struct mpcb_sim_out``."""
    _fields_ = [
        ("i_term", C.c_void_p), ("is_success", C.c_void_p), ("final_dist", C.c_void_p),
        ("x_true", C.c_void_p), ("x_est", C.c_void_p), ("ctrl", C.c_void_p),
        ("ctrlr_seq", C.c_void_p), ("status", C.c_void_p), ("iters", C.c_void_p), ("u_raw", C.c_void_p),
        ("ukf_clamped", C.c_void_p),
        ("x_true_sub", C.c_void_p), ("ctrl_sub", C.c_void_p), ("ctrlr_sub", C.c_void_p),
        ("rho", C.c_void_p),
    ]


class MpcbCounters(C.Structure):
    """``struct mpcb_counters``."""
    _fields_ = [
        ("qp_solves", C.c_int64), ("admm_iterations", C.c_int64), ("kernel_launches", C.c_int64),
        ("admm_launches", C.c_int64), ("rounds", C.c_int64), ("flip_lanes", C.c_int64),
        ("operator_rebuilds", C.c_int64),
        ("admm_ms", C.c_double), ("total_ms", C.c_double),
    ]


# every symbol include/mpcb.h declares: (restype, argtypes)
SYMBOLS = {
    "mpcb_abi_version": (C.c_int, []),
    "mpcb_last_error": (C.c_char_p, []),
    "mpcb_create": (C.c_int, [C.POINTER(MpcbProblem), C.c_int, C.POINTER(C.c_void_p)]),
    "mpcb_destroy": (C.c_int, [C.c_void_p]),
    "mpcb_batch_alloc": (C.c_int, [C.c_void_p, C.c_int64]),
    "mpcb_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "mpcb_get_counters": (C.c_int, [C.c_void_p, C.POINTER(MpcbCounters)]),
    "mpcb_stream": (C.c_void_p, [C.c_void_p]),
    "mpcb_solver_blocks": (C.c_int, [C.c_void_p]),
    "mpcb_wait_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mpcb_qp_solve": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "mpcb_qp_get_state": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mpcb_ukf_step": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "mpcb_plant_lin_step": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "mpcb_plant_rk4": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int]),
    "mpcb_simulate_discrete": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.POINTER(MpcbSimOut), C.c_int]),
    "mpcb_simulate_continuous": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_void_p,
                                           C.c_int32, C.c_int32, C.POINTER(MpcbSimOut), C.c_int]),
    "mpcb_stats": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int]),
    "mpcb_allreduce_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mpcb_noise_fill": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_double, C.c_uint64, C.c_uint64, C.c_void_p,
                                  C.c_void_p, C.c_int]),
    "mpcb_measure_fp64_peak": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double)]),
}

ABI_VERSION = 4
NSTATS = 10
_lib = None


class MpcbError(RuntimeError):
    pass


def load():
    """Load ``libmpcb.so`` and bind every symbol.  Loading needs the CUDA runtime but no GPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MpcbError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)       # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.mpcb_abi_version() != ABI_VERSION:
        raise MpcbError(f"libmpcb ABI {lib.mpcb_abi_version()} != binding ABI {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().mpcb_last_error()
        raise MpcbError(f"libmpcb error {rc}: {msg.decode() if msg else ''}")
