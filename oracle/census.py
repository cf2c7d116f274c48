"""Parity census of a batched horizon selection against the CPU oracle -- TEST INFRASTRUCTURE ONLY
(used by tests/, bench.py's cpu_baseline leg and tools/; never by the product).

SURVEY.md s.9, consequence 2: the augmented-form curve J(T) is rank-deficient up to rho_reg = 1e-12, so any two fp64
implementations -- the reference under two BLAS builds included -- differ by ~1e-7 on the selection window, and an
instance whose two best horizons are closer than that noise has no implementation-independent T*.  The census therefore
reports, per instance and aggregated (median / p99 / max), THREE relative distances

    |gpu - oracle|     the checked implementation against the fp64 oracle (plain-C restatement of the reference)
    |oracle - fp80|    the fp64 oracle against the same algorithm in x87 extended precision ("truth" of the jittered algorithm)
    |gpu - fp80|       the checked implementation against that truth

at T* and over the window [T_min, T_max], plus the argmin gap (relative distance between the best and the second-best
horizon), and classifies every instance whose gap is below 10 x the noise at its two candidates as ILL-POSED.  A T*
mismatch is *explained* iff the instance is ill-posed by that rule; unexplained mismatches are parity failures.
"""
from __future__ import annotations

import numpy as np

from . import ilqr_timeopt_batch, select_from_x0_batch


def _rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def _stats(v):
    v = np.asarray(v, dtype=float)
    if v.size == 0:
        return None
    return {"median": float(np.median(v)), "p99": float(np.percentile(v, 99)), "max": float(v.max())}


def argmin_gap(J, T, T_min, T_max):
    """Relative gap between the minimum of the window and the best OTHER horizon, per instance."""
    W = np.array(J[:, T_min - 1:T_max], dtype=float, copy=True)
    rows = np.arange(W.shape[0])
    best = W[rows, T - T_min]
    W[rows, T - T_min] = np.inf
    return (W.min(axis=1) - best) / np.abs(best), W.argmin(axis=1) + T_min


def census_from_x0(case, x0, J_gpu, T_gpu, nthreads=1, fp80_stride=8, max_listed=16, noise_factor=10.0):
    """x0 [B,n]; J_gpu [B,T_max], T_gpu [B] = the checked implementation's output for the x0 -> T* pipeline of `case`
    (U = tile(u_ref), case goal / weight).  Runs the fp64 oracle on ALL instances and the fp80 sweep on every
    `fp80_stride`-th instance plus every mismatch and every near-tie.  Returns a JSON-able dict."""
    F, _x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx, _ = case
    T_max = int(min(T_max, N))
    x0 = np.ascontiguousarray(x0, dtype=float)
    B = x0.shape[0]
    U = np.tile(np.asarray(u_ref, dtype=float).reshape(1, -1), (N, 1))
    args = (F.hop_sys, F.hop_params, N, T_min, T_max)
    kw = dict(wrap_idx=wrap_idx, nthreads=nthreads)
    Jo, To, sto = select_from_x0_batch(*args, x0, U, xg, u_ref, Q, R, alpha, w, **kw)
    J_gpu = np.asarray(J_gpu, dtype=float)
    T_gpu = np.asarray(T_gpu).astype(np.int64)
    To = To.astype(np.int64)
    ok = (sto == 0)
    rows = np.arange(B)
    win = slice(T_min - 1, T_max)
    d_go_win = _rel(J_gpu[:, win], Jo[:, win]).max(axis=1)
    d_go_star = _rel(J_gpu[rows, To - 1], Jo[rows, To - 1])
    gap_o, second_o = argmin_gap(Jo, To, T_min, T_max)
    mism = np.nonzero((T_gpu != To) & ok)[0]
    # candidates for the fp80 pass: a regular subsample + mismatches + instances whose gap is within 100 x the fp64
    # distance between the two implementations at T* (a superset of what can turn out ill-posed)
    near = np.nonzero(ok & (gap_o < 100.0 * np.maximum(d_go_star, 1e-12)))[0]
    sub = np.arange(0, B, max(1, int(fp80_stride)))
    idx = np.unique(np.concatenate([sub, mism, near]))
    J80, T80, st80 = select_from_x0_batch(*args, x0[idx], U, xg, u_ref, Q, R, alpha, w, f80=True, **kw)
    T80 = T80.astype(np.int64)
    r = np.arange(idx.size)
    d_o8_win = _rel(Jo[idx][:, win], J80[:, win]).max(axis=1)
    d_g8_win = _rel(J_gpu[idx][:, win], J80[:, win]).max(axis=1)
    d_o8_star = _rel(Jo[idx, To[idx] - 1], J80[r, To[idx] - 1])
    d_g8_star = _rel(J_gpu[idx, To[idx] - 1], J80[r, To[idx] - 1])
    gap80, second80 = argmin_gap(J80, T80, T_min, T_max)
    # ill-posedness (SURVEY s.9): gap between the two best horizons of the TRUTH curve below noise_factor x the largest
    # fp64-vs-truth distance of either implementation at those two horizons
    cand_a, cand_b = T80, second80
    noise = np.maximum.reduce([
        _rel(Jo[idx, cand_a - 1], J80[r, cand_a - 1]), _rel(Jo[idx, cand_b - 1], J80[r, cand_b - 1]),
        _rel(J_gpu[idx, cand_a - 1], J80[r, cand_a - 1]), _rel(J_gpu[idx, cand_b - 1], J80[r, cand_b - 1])])
    ill = gap80 < noise_factor * noise
    pos = {int(b): i for i, b in enumerate(idx)}
    listed, unexplained = [], []
    for b in mism:
        i = pos[int(b)]
        tg, to = int(T_gpu[b]), int(To[b])
        pair_gap = float(abs(J80[i, tg - 1] - J80[i, to - 1]) / abs(J80[i, to - 1]))
        pair_noise = float(max(_rel(Jo[b, tg - 1], J80[i, tg - 1]), _rel(Jo[b, to - 1], J80[i, to - 1]),
                               _rel(J_gpu[b, tg - 1], J80[i, tg - 1]), _rel(J_gpu[b, to - 1], J80[i, to - 1])))
        explained = pair_gap < noise_factor * pair_noise
        rec = {"instance": int(b), "T_gpu": tg, "T_oracle": to, "T_fp80": int(T80[i]), "gap_between_candidates_fp80": pair_gap,
               "noise_at_candidates": pair_noise, "ill_posed": bool(explained)}
        (listed if explained else unexplained).append(rec)
    in_sub = np.isin(idx, sub)
    return {
        "checked": int(B), "oracle_errors": int((~ok).sum()),
        "T_star_mismatches": int(mism.size), "T_star_mismatches_unexplained": len(unexplained),
        "mismatch_rate": float(mism.size) / max(1, B),
        "T_star_vs_fp80_on_sample": {"sample": int(in_sub.sum()),
                                     "gpu_differs": int((T_gpu[idx][in_sub] != T80[in_sub]).sum()),
                                     "oracle_differs": int((To[idx][in_sub] != T80[in_sub]).sum())},
        "rel_J_at_Tstar": {"gpu_vs_oracle": _stats(d_go_star[ok]), "oracle_vs_fp80": _stats(d_o8_star[in_sub]),
                           "gpu_vs_fp80": _stats(d_g8_star[in_sub])},
        "rel_J_window": {"gpu_vs_oracle": _stats(d_go_win[ok]), "oracle_vs_fp80": _stats(d_o8_win[in_sub]),
                         "gpu_vs_fp80": _stats(d_g8_win[in_sub])},
        "argmin_gap_oracle": {"min": float(gap_o[ok].min()), "p1": float(np.percentile(gap_o[ok], 1)),
                              "median": float(np.median(gap_o[ok]))},
        "ill_posed": {"rule": f"fp80 gap between the two best horizons < {noise_factor:g} x max fp64-vs-fp80 distance "
                              "(either implementation) at those horizons",
                      "count_in_fp80_pass": int(ill.sum()), "fp80_pass": int(idx.size),
                      "count_on_regular_sample": int(ill[in_sub].sum()), "regular_sample": int(in_sub.sum()),
                      "instances": [int(b) for b in idx[ill][:max_listed]]},
        "mismatch_detail": (unexplained + listed)[:max_listed],
    }


# ---------------------------------------------------------------------------------------------------------------------
# iterated solves (HOP-DDP): which instances have a T_hist that the reference computation itself reproduces?
# ---------------------------------------------------------------------------------------------------------------------
PERTURBATIONS = ("fp80_selection", "x0+1e-15", "x0-1e-15", "x0*(1+4e-16)", "x0*(1-4e-16)", "x0+3e-15")


def same_history(p, q, b):
    """T_hist of instance b identical in two result dicts (n_hist, T_hist)."""
    return p["n_hist"][b] == q["n_hist"][b] and np.array_equal(p["T_hist"][b, :p["n_hist"][b]], q["T_hist"][b, :q["n_hist"][b]])


def ddp_oracle_census(case, N, x0s, max_iter=12, nthreads=1):
    """The fp64 oracle and six rounding-level perturbations of it on the same initial states (the selection sweep in x87
    extended precision; x0 +- 1e-15; x0 (1 +- 4e-16); x0 + 3e-15).  Returns (oracle result, well [B] bool = T_hist identical
    in all seven runs, {perturbation: fraction of instances whose T_hist it flips}, seconds of the unperturbed run)."""
    import time
    F, _x0, xg, u_ref, Q, R, alpha, w, _N, T_min, T_max, wrap_idx, _ = case
    T_max = min(T_max, N)
    x0s = np.ascontiguousarray(x0s, dtype=float)
    kw = dict(max_iter=max_iter, use_central_diff=False, nthreads=nthreads)
    a = lambda x: (F.hop_sys, F.hop_params, N, T_min, T_max, x, np.tile(u_ref, (N, 1)), xg, u_ref, Q, R, alpha, w, wrap_idx)  # noqa: E731
    t0 = time.perf_counter()
    o = ilqr_timeopt_batch(*a(x0s), **kw)
    dt = time.perf_counter() - t0
    variants = {"fp80_selection": ilqr_timeopt_batch(*a(x0s), f80_select=True, **kw),
                "x0+1e-15": ilqr_timeopt_batch(*a(x0s + 1e-15), **kw), "x0-1e-15": ilqr_timeopt_batch(*a(x0s - 1e-15), **kw),
                "x0*(1+4e-16)": ilqr_timeopt_batch(*a(x0s * (1.0 + 4e-16)), **kw),
                "x0*(1-4e-16)": ilqr_timeopt_batch(*a(x0s * (1.0 - 4e-16)), **kw),
                "x0+3e-15": ilqr_timeopt_batch(*a(x0s + 3e-15), **kw)}
    B = len(x0s)
    stable = {name: np.array([same_history(o, v, b) for b in range(B)]) for name, v in variants.items()}
    well = np.logical_and.reduce(list(stable.values()))
    return o, well, {name: float(1.0 - s_.mean()) for name, s_ in stable.items()}, dt
