"""Full-horizon parity report: the CUDA engine against the oracle, lane by lane (ORACLE -- test infrastructure).

Only ``tests/``, ``tools/parity_report.py`` and ``bench.py``'s parity slice (outside every timed region) import this.

The closed loop follows reference ``src/trajectorySimulate.py:285-356`` for ``nsim`` control steps.  Both sides run the
same float64 algorithm and differ in summation order only, so the discrete record of a lane -- ADMM iteration count,
OSQP status and controller choice of every solve up to ``i_term``, and ``i_term`` itself -- is either EXACTLY equal or
diverges at some first step (a termination test that falls the other way on a 1e-13 difference, typically once adaptive
rho has been driven to 1e2..1e6 by a run of infeasible QPs and the KKT system has condition 1e9+).  The report says how
many lanes are exact over the whole horizon, where and at which rho the others diverge, how far the controls are apart
on the matching prefix (bar: 1e-4 abs, BASELINE.json north_star), and how far the diverged lanes end up from the
oracle's (they are the same controller on a perturbed path: the final-distance statistics must agree).
"""
import numpy as np


def _host(a):
    if a is None:
        return None
    if hasattr(a, "cpu"):
        a = a.cpu().numpy()
    return np.asarray(a)


RHO_DECISION_TOL = 1e-2
RHO_SAME_TOL = 1e-6


def full_horizon_report(got, ref, u_bar=1e-4, rho_tol=RHO_DECISION_TOL):
    """``got``: BatchSimRun (engine layout ``[field, T, B]``); ``ref``: dict of ``simulate_discrete_batch``
    (``[T, B, field]``).  Returns a JSON-serialisable dict."""
    it_g, it_r = _host(got.iters).astype(np.int64), np.asarray(ref["iters"], np.int64)            # [nsim, B]
    st_g, st_r = _host(got.status).astype(np.int64), np.asarray(ref["status"], np.int64)
    cs_g, cs_r = _host(got.ctrlr_seq).astype(np.int64), np.asarray(ref["ctrlr_seq"], np.int64)
    term_g, term_r = _host(got.i_term).astype(np.int64), np.asarray(ref["i_term"], np.int64)
    u_g = _host(got.ctrl_hist).transpose(1, 2, 0)                                                  # [T1, B, 2]
    u_r = np.asarray(ref["ctrl_hist"])
    x_g = _host(got.x_true).transpose(1, 2, 0)
    x_r = np.asarray(ref["x_true"])
    nsim, B = it_r.shape
    steps = np.arange(nsim)[:, None]
    live = steps < np.minimum(term_g, term_r)[None, :]
    differ = ((it_g != it_r) | (st_g != st_r) | (cs_g != cs_r)) & live
    rho_g = _host(getattr(got, "rho", None))
    rho_r = ref.get("rho_hist")
    differ_strict = differ.copy()
    if rho_g is not None and rho_r is not None:
        # adaptive rho is a discrete decision too (OSQP adapts when the estimate leaves [rho/5, 5 rho]): a lane where one side
        # adapted and the other did not has diverged even while the iteration counts still coincide
        # (RHO_DECISION_TOL: an adaptation multiplies rho by >= 5 or <= 1/5, so 1 % separates "adapted differently" from the
        # 1e-9..1e-4 by which two float64 evaluations of sqrt(pri/dua) differ once the KKT system is ill-conditioned)
        with np.errstate(invalid="ignore"):
            gap = np.abs(rho_g - rho_r) / np.abs(rho_r)
        gap = np.nan_to_num(gap, nan=0.0)
        differ |= (gap > rho_tol) & live
        # ... but a rho that differs in the 5th digit already makes the NEXT solve a different OSQP run (same decisions, iterates
        # apart by a fraction of OSQP's own 1e-3 termination tolerance): controls are compared on the strict prefix, where rho
        # agrees to RHO_SAME_TOL as well
        differ_strict |= (gap > RHO_SAME_TOL) & live
    first = np.where(differ.any(axis=0), differ.argmax(axis=0), nsim)                               # first differing solve
    first_s = np.where(differ_strict.any(axis=0), differ_strict.argmax(axis=0), nsim)
    # a lane also diverges where the two sides stop at different steps
    first = np.where(term_g != term_r, np.minimum(first, np.minimum(term_g, term_r)), first)
    first_s = np.minimum(first_s, first)
    exact = (first >= nsim) & (term_g == term_r)
    exact_s = (first_s >= nsim) & (term_g == term_r)
    # controls / states on the matching prefix: step i's command is ctrl[i+1]; it is comparable while solves 0..i matched
    upto = np.minimum(first, np.minimum(term_g, term_r))                                           # solves 0..upto-1 match
    upto_s = np.minimum(first_s, np.minimum(term_g, term_r))
    tt = np.arange(nsim + 1)[:, None]
    pre = tt <= upto_s[None, :]
    du = np.where(pre[:, :, None], np.abs(u_g - u_r), 0.0)
    dx = np.where(pre[:, :, None], np.abs(x_g - x_r), 0.0)
    du_dec = np.nan_to_num(np.where((tt <= upto[None, :])[:, :, None], np.abs(u_g - u_r), 0.0), nan=0.0)
    du = np.nan_to_num(du, nan=0.0)
    dx = np.nan_to_num(dx, nan=0.0)
    rep = {
        "lanes": int(B), "steps": int(nsim),
        "exact_lanes": int(exact.sum()), "exact_frac": float(exact.mean()), "exact_strict_frac": float(exact_s.mean()),
        "solves_strict_prefix": int(np.minimum(upto_s, nsim).sum()), "max_du_decision_prefix": float(du_dec.max()),
        "solves_compared": int(live.sum()), "solves_exact_prefix": int(np.minimum(upto, nsim).sum()),
        "max_du_prefix": float(du.max()), "max_dx_prefix": float(dx.max()),
        "u_bar": u_bar, "du_within_bar": bool(du.max() <= u_bar),
    }
    div = np.nonzero(~exact)[0]
    if div.size:
        rho_hist = ref.get("rho_hist")
        fs = first[div]
        rep["diverged_lanes"] = int(div.size)
        rep["first_divergence_step"] = {"min": int(fs.min()), "median": float(np.median(fs)), "max": int(fs.max())}
        if rho_hist is not None:
            # rho the oracle's lane carried INTO the first differing solve (= rho after the previous one)
            r_at = np.array([rho_hist[max(int(s) - 1, 0), b] if s < nsim else np.nan for s, b in zip(fs, div)])
            r_at = r_at[np.isfinite(r_at)]
            if r_at.size:
                rep["rho_at_divergence"] = {"min": float(r_at.min()), "median": float(np.median(r_at)), "max": float(r_at.max())}
        fd_g = _host(got.final_dist)[div]
        xr_fin = np.array([x_r[max(int(term_r[b]) - 1, 0), b] for b in div])
        # the oracle's final distance: ||x_true[i_term-1] - xr|| is what the engine reports; recompute both from telemetry
        xg_fin = np.array([x_g[max(int(term_g[b]) - 1, 0), b] for b in div])
        d_fin = np.linalg.norm(xg_fin[:, :2] - xr_fin[:, :2], axis=1)
        rep["diverged_final_pos_gap"] = {"median": float(np.median(d_fin)), "max": float(d_fin.max())}
        rep["diverged_i_term_gap_max"] = int(np.abs(term_g[div] - term_r[div]).max())
        # status at the first differing solve, both sides
        pairs = {}
        for s, b in zip(fs, div):
            if s < nsim:
                key = f"{int(st_r[s, b])}/{int(it_r[s, b])}->{int(st_g[s, b])}/{int(it_g[s, b])}"
                pairs[key] = pairs.get(key, 0) + 1
        rep["first_divergence_kinds"] = dict(sorted(pairs.items(), key=lambda kv: -kv[1])[:6])
    # batch statistics: the two sides are the same controller, so the Monte-Carlo outputs must agree
    fd_all_g = _host(got.final_dist)
    fin_r = np.array([x_r[max(int(term_r[b]) - 1, 0), b] for b in range(B)])
    rep["mean_final_dist_engine"] = float(np.nanmean(fd_all_g))
    rep["success_engine"] = int(_host(got.isSuccess).sum())
    rep["i_term_equal_frac"] = float((term_g == term_r).mean())
    rep["mean_final_pos_norm_oracle"] = float(np.nanmean(np.linalg.norm(fin_r[:, :2], axis=1)))
    return rep


def as_engine_layout(ref):
    """A ``simulate_discrete_batch`` / ``c_ref.simulate_discrete`` dict dressed as a BatchSimRun (``[field, T, B]``), so that
    two CPU implementations can be compared with each other by ``full_horizon_report`` (the control of the parity tests)."""
    from types import SimpleNamespace
    B = len(ref["i_term"])
    x = np.asarray(ref["x_true"])
    fin = np.array([x[max(int(t) - 1, 0), b, :2] for b, t in enumerate(ref["i_term"])])
    return SimpleNamespace(iters=ref["iters"], status=ref["status"], ctrlr_seq=ref["ctrlr_seq"], i_term=ref["i_term"],
                           ctrl_hist=np.asarray(ref["ctrl_hist"]).transpose(2, 0, 1), x_true=x.transpose(2, 0, 1),
                           rho=ref.get("rho_hist"), final_dist=ref.get("final_dist", np.linalg.norm(fin, axis=1)),
                           isSuccess=np.asarray(ref.get("isSuccess", np.zeros(B, int))))


def markdown_row(name, rep):
    d = rep.get("first_divergence_step")
    r = rep.get("rho_at_divergence")
    return (f"| {name} | {rep['lanes']} x {rep['steps']} | {rep['exact_lanes']} ({100 * rep['exact_frac']:.1f} %) | "
            f"{rep['solves_exact_prefix']} / {rep['solves_compared']} | "
            f"{(str(d['min']) + ' / ' + format(d['median'], '.0f')) if d else '-'} | "
            f"{(format(r['min'], '.3g') + ' / ' + format(r['median'], '.3g')) if r else '-'} | "
            f"{rep['max_du_prefix']:.1e} | {rep['max_dx_prefix']:.1e} | "
            f"{(format(rep['diverged_final_pos_gap']['median'], '.2e') + ' / ' + format(rep['diverged_final_pos_gap']['max'], '.2e')) if 'diverged_final_pos_gap' in rep else '-'} |")


MARKDOWN_HEADER = ("| workload | lanes x steps | lanes exact to i_term | solves on the exact prefix | first divergence step (min / median) | "
                   "rho there (min / median) | max abs du on prefix | max abs dx on prefix | final position gap of diverged lanes, m (median / max) |\n"
                   "|---|---|---|---|---|---|---|---|---|")
