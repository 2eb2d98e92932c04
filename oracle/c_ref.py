"""ctypes face of ``oracle/c/mpc_ref.c`` -- the compiled CPU twin of the reference's per-trajectory loop (ORACLE: test
infrastructure; only ``tests/``, ``__graft_entry__`` and ``bench.py``'s CPU legs import it).

``build()`` compiles it with gcc (``-O3 -march=x86-64-v3 -pthread``: AVX2 + FMA, present on the build container and the GPU box alike) into ``oracle/_build/libmpcref.so`` (git-ignored, travels to the GPU box);
``simulate_discrete`` has the argument meaning and array layouts of ``Engine.simulate_discrete`` and returns a dict shaped like
``oracle.batched_ref.simulate_discrete_batch``'s, so the same comparison code serves both.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "mpc_ref.c")
LIB = os.path.join(HERE, "_build", "libmpcref.so")
_lib = None


def build(force: bool = False) -> str:
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    hdr = os.path.join(os.path.dirname(HERE), "include", "mpcb.h")
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(f) > os.path.getmtime(LIB) for f in (SRC, hdr))
    if force or stale:
        cmd = [os.environ.get("CC", "gcc"), "-O3", "-march=x86-64-v3", "-pthread", "-shared", "-fPIC", "-std=gnu11", "-o", LIB, SRC, "-lm"]
        subprocess.check_call(cmd)
    return LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        lib = C.CDLL(LIB)
        from mpc_arpo_project_b200._lib import MpcbProblem, MpcbSimOut, ABI_VERSION
        lib.mpcref_abi_version.restype = C.c_int
        assert lib.mpcref_abi_version() == ABI_VERSION, "oracle/c/mpc_ref.c was built against another include/mpcb.h: rebuild"
        lib.mpcref_simulate_discrete.restype = C.c_int
        lib.mpcref_simulate_discrete.argtypes = [C.POINTER(MpcbProblem), C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                                 C.POINTER(MpcbSimOut), C.c_int, C.c_void_p]
        _lib = lib
    return _lib


def simulate_discrete(problem, x0, noise, nsteps, nthreads=0, record=True):
    """``x0[4, B]``, ``noise[R, 2, B]`` or None -> dict with the keys of ``simulate_discrete_batch`` (``[T, B, field]``)."""
    from mpc_arpo_project_b200.engine import fill_problem_struct
    from mpc_arpo_project_b200._lib import MpcbSimOut
    lib = load()
    cp, keep = fill_problem_struct(problem)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    B = x0.shape[1]
    T1 = nsteps + 1
    R = 0 if noise is None else noise.shape[0]
    noise = None if noise is None else np.ascontiguousarray(noise, dtype=np.float64)
    out = MpcbSimOut()
    arr = {"i_term": np.zeros(B, np.int32), "is_success": np.zeros(B, np.int32), "final_dist": np.zeros(B), "ukf_clamped": np.zeros(B, np.int32)}
    if record:
        arr.update(x_true=np.full((4, T1, B), np.nan), x_est=np.full((6, T1, B), np.nan), ctrl=np.full((2, T1, B), np.nan),
                   ctrlr_seq=np.zeros((T1 - 1, B), np.uint8), status=np.zeros((T1 - 1, B), np.int8), iters=np.zeros((T1 - 1, B), np.int16),
                   u_raw=np.full((2, T1 - 1, B), np.nan), rho=np.full((T1 - 1, B), np.nan))
    for k, a in arr.items():
        setattr(out, k, a.ctypes.data)
    counts = np.zeros(2, np.int64)
    rc = lib.mpcref_simulate_discrete(C.byref(cp), B, int(nsteps), x0.ctypes.data, noise.ctypes.data if noise is not None else None, R,
                                      C.byref(out), int(nthreads), counts.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"mpcref_simulate_discrete failed with {rc}")
    res = dict(i_term=arr["i_term"].astype(np.int64), isSuccess=arr["is_success"], final_dist=arr["final_dist"],
               ukf_clamped=arr["ukf_clamped"].astype(bool), qp_solves=int(counts[0]), admm_iterations=int(counts[1]))
    if record:
        res.update(x_true=arr["x_true"].transpose(1, 2, 0), x_est=arr["x_est"].transpose(1, 2, 0), ctrl_hist=arr["ctrl"].transpose(1, 2, 0),
                   u_raw=arr["u_raw"].transpose(1, 2, 0), ctrlr_seq=arr["ctrlr_seq"], status=arr["status"].astype(int),
                   iters=arr["iters"].astype(int), rho_hist=arr["rho"])
    return res
