"""Generate ``tests/golden/*.npz`` by running the REFERENCE'S OWN driver code.

ORACLE tooling (test infrastructure).  Runs only in the build container, where
``/root/reference`` exists; the committed fixtures travel to the GPU box.

What this pins and what it cannot:
  * The reference's ``src/trajectorySimulate.py``, ``src/trajectorySimulateC.py``,
    ``src/simhelpers.py`` and ``src/mpcsim.py`` are imported UNMODIFIED from a scratch copy
    (one line patched: ``simhelpers.py:66-67`` indexes a tuple with ``numpy.bool_``,
    which NumPy >= 2.3 rejects -- the patch wraps it in ``int()``).
  * The third-party packages they import are absent from this image (``osqp``,
    ``filterpy``, ``control``, ``matplotlib``); thin shims route those calls to the
    restatements in ``oracle/osqp_ref.py`` / ``ukf_ref.py`` / ``control_ref.py``.
  => The fixtures pin the reference's QP assembly (P, q, A, l, u at setup and every
     ``prob.update`` payload) and its closed-loop driver logic exactly, *given* the
     restated third-party numerics.  They do NOT pin OSQP/filterpy/control themselves
     (PARITY UNPINNED at that boundary; see oracle/__init__.py).

Usage:  python oracle/gen_golden.py  [--out tests/golden]
"""
import argparse
import os
import shutil
import sys
import tempfile
import types

import numpy as np
from scipy import sparse

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.osqp_ref import OSQPRef                     # noqa: E402
from oracle.ukf_ref import UKFRef, MerweScaledSigmaPointsRef   # noqa: E402
from oracle import control_ref                          # noqa: E402

CAPTURE = {}


class _OSQPShim:
    """osqp.OSQP protocol -> OSQPRef, recording every payload the reference sends."""

    def setup(self, P, q, A, l, u, **kw):
        self._Acsc = sparse.csc_matrix(A)
        self._Acsc.sort_indices()
        self._ref = OSQPRef()
        self._ref.setup(P, q, self._Acsc, l, u, **kw)
        CAPTURE.setdefault('setup', []).append(dict(P=sparse.csc_matrix(P).toarray(), q=np.array(q), A=self._Acsc.toarray(),
                                                    l=np.array(l), u=np.array(u), kw=dict(kw)))
        CAPTURE['updates'] = []
        CAPTURE['solves'] = []

    def solve(self):
        r = self._ref.solve()
        CAPTURE['solves'].append(dict(status=r.info.status, iter=r.info.iter, x=np.array(r.x)))
        return r

    def update(self, l=None, u=None, Ax=None, **kw):
        assert not kw, kw
        A = None
        if Ax is not None:
            # OSQP interprets Ax as the value array of the CSC matrix given at setup
            M = self._Acsc.copy()
            assert len(Ax) == M.nnz, (len(Ax), M.nnz)
            M.data = np.array(Ax, float)
            A = M.toarray()
        CAPTURE['updates'].append(dict(l=None if l is None else np.array(l), u=None if u is None else np.array(u), A=A))
        self._ref.update(l=l, u=u, A=A)


class _UKFShim(UKFRef):
    def __init__(self, dim_x, dim_z, dt, fx, hx, points):
        super().__init__(dim_x, dim_z, fx, hx, points)

    def update(self, z):
        super().update(z)


def _install_shims():
    osqp = types.ModuleType('osqp')
    osqp.OSQP = _OSQPShim
    sys.modules['osqp'] = osqp
    fp = types.ModuleType('filterpy')
    fk = types.ModuleType('filterpy.kalman')
    fk.UnscentedKalmanFilter = _UKFShim
    fk.MerweScaledSigmaPoints = lambda n, alpha, beta, kappa: MerweScaledSigmaPointsRef(n, alpha, beta, kappa)
    fp.kalman = fk
    sys.modules['filterpy'] = fp
    sys.modules['filterpy.kalman'] = fk
    ct = types.ModuleType('control')
    ct.dlqr = lambda A, B, Q, R, integral_action=None: (control_ref.dlqr_integral(A, B, Q, R, integral_action), None, None)
    ct.acker = control_ref.acker
    ct.white_noise = lambda T, Q, dt=0: control_ref.white_noise(T, Q, lambda size: np.random.normal(0, 1, size))
    sys.modules['control'] = ct
    mpl = types.ModuleType('matplotlib')
    plt = types.ModuleType('matplotlib.pyplot')
    lines = types.ModuleType('matplotlib.lines')
    lines.Line2D = object
    mpl.pyplot, mpl.lines = plt, lines
    sys.modules['matplotlib'] = mpl
    sys.modules['matplotlib.pyplot'] = plt
    sys.modules['matplotlib.lines'] = lines


def import_reference(ref_root='/root/reference'):
    """Copy ``src/`` to a scratch dir, apply the one NumPy-compat patch, import it."""
    tmp = tempfile.mkdtemp(prefix='refcopy_')
    shutil.copytree(os.path.join(ref_root, 'src'), os.path.join(tmp, 'src'))
    p = os.path.join(tmp, 'src', 'simhelpers.py')
    txt = open(p).read()
    a = "C1 = (-1, 1)[xest[2] >= 0]"
    b = "C2 = (-1, 1)[xest[3] >= 0]"
    assert a in txt and b in txt
    txt = txt.replace(a, "C1 = (-1, 1)[int(xest[2] >= 0)]").replace(b, "C2 = (-1, 1)[int(xest[3] >= 0)]")
    open(p, 'w').write(txt)
    _install_shims()
    sys.path.insert(0, tmp)
    import src.mpcsim as ref_mpcsim
    import src.trajectorySimulate as ref_ts
    import src.trajectorySimulateC as ref_tsc
    return ref_mpcsim, ref_ts.trajectorySimulate, ref_tsc.trajectorySimulateC


# ------------------------------------------------------------------ parameter sets
def make_params(M, case):
    """Parameter sets after test/traj_eval_radial.py:17-72, traj_eval_radialC.py:17-75,
    traj_eval_in_track.py:14-66 (u_lim supplied), disturbRejComp.py:17-72.  ``M`` is a
    module exposing the mpcsim classes (the reference's or the product's).  The literals live in the
    package (``mpc_arpo_project_b200/presets.py``) so that product code never imports the oracle."""
    from mpc_arpo_project_b200.presets import make_params as _mk
    return _mk(case, M)


CASES = {
    # name: (simulator, params)
    'radial_nx10_noise': ('D', dict(Nx=10, sigma=0.75, noise_length=50, T_final=30)),
    'radial_nx10_quiet': ('D', dict(Nx=10, sigma=0.1, noise_length=50, T_final=40)),
    'radial_nx40_nonoise': ('D', dict(Nx=40, sigma=None, T_final=20)),
    'radial_nx40_debris': ('D', dict(Nx=40, sigma=0.75, noise_length=50, T_final=20, debris=((40., 0.), 5., 20))),
    'intrack_dv_nx20': ('D', dict(Nx=20, inTrack=True, isDeltaV=True, isReject=False, sigma=None, T_final=30)),
    'radial_nx30_norej': ('D', dict(Nx=30, sigma=0.7, noise_length=20, isReject=False, T_final=25)),
    'contC_nx10_accel': ('C', dict(Nx=10, sigma=0.0012, noise_length=4, T_cont=0.001, T_final=4)),
    'contC_nx10_dv': ('C', dict(Nx=10, sigma=0.0012, noise_length=4, T_cont=0.001, T_final=3, isDeltaV=True)),
}


def run_case(name, refmods):
    ref_mpcsim, ref_ts, ref_tsc = refmods
    kind, case = CASES[name]
    sc, mp, fp, debris = make_params(ref_mpcsim, case)
    CAPTURE.clear()
    if kind == 'D':
        run = ref_ts(sc, mp, fp, debris)            # seeds numpy RNG with 123 itself
    else:
        np.random.seed(321)                         # trajectorySimulateC leaves the RNG unseeded (:28)
        run = ref_tsc(sc, mp, fp, debris)
    st = CAPTURE['setup'][0]
    ups = CAPTURE['updates']
    nup = min(len(ups), 24)
    out = dict(
        P=st['P'], q=st['q'], A=st['A'], l=st['l'], u=st['u'],
        upd_l=np.array([ups[k]['l'] for k in range(nup)]), upd_u=np.array([ups[k]['u'] for k in range(nup)]),
        upd_A_last=ups[nup - 1]['A'] if nup and ups[nup - 1]['A'] is not None else np.zeros(0),
        solve_status=np.array([s['status'] for s in CAPTURE['solves']]),
        solve_iter=np.array([s['iter'] for s in CAPTURE['solves']]),
        i_term=run.i_term, isSuccess=run.isSuccess, x_true_pcw=np.array(run.x_true_pcw), x_est=np.array(run.x_est),
        ctrl_hist=np.array(run.ctrl_hist), ctrlr_seq=np.array(run.ctrlr_seq), noise_hist=np.array(run.noise_hist),
    )
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(ROOT, 'tests', 'golden'))
    ap.add_argument('--only', default=None)
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    refmods = import_reference()
    for name in CASES:
        if args.only and name != args.only:
            continue
        out = run_case(name, refmods)
        if CASES[name][0] == 'C':
            # continuous telemetry is per 1 ms substep; keep it small
            it = int(out['i_term'])
            out['x_true_pcw'] = out['x_true_pcw'][:, :it:50]
            out['ctrl_hist'] = out['ctrl_hist'][:, :it:50]
            out['ctrlr_seq'] = out['ctrlr_seq'][:it:50]
        path = os.path.join(args.out, f'ref_{name}.npz')
        np.savez_compressed(path, **out)
        print(f'{name}: i_term={out["i_term"]} success={out["isSuccess"]} solves={len(out["solve_iter"])} '
              f'-> {path} ({os.path.getsize(path) / 1024:.0f} KiB)')


if __name__ == '__main__':
    main()
