# -*- coding: utf-8 -*-
"""Drop-in for the reference's run_suite.py: same case registry, trial sampling, CLI flags and CSV schema
(run_suite.py:55-81,88-240), with all trials of a case solved in ONE batched device call.

Usage:  python run_suite.py --cases Segway_Balance --trials 25 --max-iter 12
Only the "ourmethod" solver (HOP) runs on the B200; baseline1/baseline2 are the reference's CPU comparators.
`total_time` is the batch's device time (sum of the four phase timers) divided by the number of trials."""
from __future__ import annotations

import argparse
import os
from typing import Callable, Dict, List, Tuple

import numpy as np

from _bridge import api, dev
from solver import ilqr_timeopt_baseline1, ilqr_timeopt_baseline2, ilqr_timeopt_ourmethod
from systems import make_cartpole_swingup, make_double_integrator, make_quadrotor, make_segway_balance
from utils import wrap_error


def _rng(seed: int) -> np.random.Generator:
    return np.random.default_rng(int(seed))


def sample_x(base: np.ndarray, sigma: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    base = np.asarray(base, dtype=float).reshape(-1)
    sigma = np.asarray(sigma, dtype=float).reshape(-1)
    if sigma.size == 1:
        sigma = np.full_like(base, float(sigma))
    return base + sigma * rng.standard_normal(base.shape)


CaseMaker = Callable[[], Tuple]

CASES: List[Tuple[str, CaseMaker, Dict[str, np.ndarray]]] = [
    ("DoubleIntegrator", make_double_integrator, dict(sigma_x0=np.array([0.2, 0.2]), sigma_xg=np.array([0.0, 0.0]))),
    ("Cartpole_SwingUp", make_cartpole_swingup, dict(sigma_x0=np.zeros(4), sigma_xg=np.zeros(4))),
    ("Quadrotor", make_quadrotor, dict(sigma_x0=np.array([0.4, 0.4, 0.4] + [0.0] * 9), sigma_xg=np.zeros(12))),
    ("Segway_Balance", make_segway_balance, dict(sigma_x0=np.full(4, 0.02), sigma_xg=np.zeros(4))),
]

SOLVERS = {"ourmethod": ilqr_timeopt_ourmethod, "baseline1": ilqr_timeopt_baseline1, "baseline2": ilqr_timeopt_baseline2}


def sample_trials(case_name, x0_base, xg_base, sigmas, trials, seed):
    """Trial 0 is the nominal case; later trials follow run_suite.py:108-120 (x0 then xg from one generator)."""
    rng = _rng(seed + hash(case_name) % 10_000)
    x0s, xgs = [], []
    for trial in range(int(trials)):
        if trial == 0:
            x0s.append(np.asarray(x0_base, float).reshape(-1)); xgs.append(np.asarray(xg_base, float).reshape(-1))
        else:
            x0s.append(sample_x(x0_base, sigmas["sigma_x0"], rng)); xgs.append(sample_x(xg_base, sigmas["sigma_xg"], rng))
    return np.stack(x0s), np.stack(xgs)


def run_case(case_name, maker, sigmas, *, outdir, trials, seed, solvers, max_iter, S_window, use_central_diff, success_tol):
    import pandas as pd
    case = maker()
    F, x0_base, xg_base, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx, extra = case
    case_dir = os.path.join(outdir, case_name)
    os.makedirs(case_dir, exist_ok=True)
    x0s, xgs = sample_trials(case_name, x0_base, xg_base, sigmas, trials, seed)
    rows = []
    for solver_name in solvers:
        if solver_name != "ourmethod":
            raise NotImplementedError(f"solver {solver_name!r} is a CPU baseline of the reference; the B200 suite runs 'ourmethod'")
        r = api.ilqr_timeopt_batched(case, dev(x0s), xg=dev(xgs), max_iter=max_iter, use_central_diff=use_central_diff)
        st = r["status"].cpu().numpy() & 0xFF
        nh = r["n_hist"].cpu().numpy(); Jh = r["J_hist"].cpu().numpy(); Ts = r["T_star"].cpu().numpy()
        X = r["X"].cpu().numpy()
        per_solve = float(sum(r["timers"].values())) / max(1, len(x0s))
        for trial in range(len(x0s)):
            if st[trial]:
                rows.append(dict(case=case_name, trial=trial, solver=solver_name, status="crash", T_star=int(T_min), J_star=float("nan"),
                                 total_time=per_solve, final_err=float("nan"), success=False, n_iter=0,
                                 solver_error="FloatingPointError" if st[trial] == 1 else "LinAlgError"))
                continue
            J_star = float(Jh[trial, nh[trial] - 1]) if nh[trial] else float("inf")
            eT = wrap_error(X[trial, Ts[trial]] - xgs[trial], wrap_idx)
            final_err = float(np.linalg.norm(np.asarray(eT, float).reshape(-1)))
            success = bool(np.isfinite(J_star) and np.isfinite(final_err) and final_err <= float(success_tol))
            rows.append(dict(case=case_name, trial=trial, solver=solver_name, status="ok" if success else "fail", T_star=int(Ts[trial]),
                             J_star=J_star, total_time=per_solve, final_err=final_err, success=success, n_iter=int(nh[trial]),
                             solver_error=None))
    df = pd.DataFrame(rows)
    df["best_J"] = df.groupby(["case", "trial"])["J_star"].transform("min")
    df["cost_ratio_best"] = df["J_star"] / df["best_J"]
    df["time_base"] = np.nan
    df["time_ratio_base"] = np.nan
    df.to_csv(os.path.join(case_dir, "summary_all.csv"), index=False)
    _aggregate(df).to_csv(os.path.join(case_dir, "summary_agg.csv"), index=False)
    return df


def _aggregate(df):
    return (df.groupby(["case", "solver"])
              .agg(n=("trial", "count"), success_rate=("success", "mean"), T_median=("T_star", "median"),
                   J_median=("J_star", "median"), time_median=("total_time", "median"),
                   ratio_cost_median=("cost_ratio_best", "median"), ratio_time_median=("time_ratio_base", "median"))
              .reset_index())


def main():
    import pandas as pd
    ap = argparse.ArgumentParser()
    ap.add_argument("--outdir", type=str, default="ilqr_results", help="output directory")
    ap.add_argument("--trials", type=int, default=25, help="trials per case (same for all cases)")
    ap.add_argument("--seed", type=int, default=0, help="random seed")
    ap.add_argument("--max-iter", type=int, default=12, help="max iLQR iterations per run")
    ap.add_argument("--S-window", type=int, default=20, help="onepass search half-window (unused by ourmethod)")
    ap.add_argument("--use-central-diff", action="store_true", help="use central differences for linearization")
    ap.add_argument("--success-tol", type=float, default=0.5, help="terminal error norm threshold for success")
    ap.add_argument("--solvers", type=str, default="ourmethod", help="comma-separated subset (B200: ourmethod)")
    ap.add_argument("--cases", type=str, default="", help="comma-separated case names (default: all)")
    args = ap.parse_args()
    os.makedirs(args.outdir, exist_ok=True)
    solvers = [s.strip() for s in args.solvers.split(",") if s.strip()]
    for s in solvers:
        if s not in SOLVERS:
            raise ValueError(f"Unknown solver: {s}. Options: {list(SOLVERS)}")
    if args.cases.strip():
        wanted = set(c.strip() for c in args.cases.split(",") if c.strip())
        cases_sel = [c for c in CASES if c[0] in wanted]
        if not cases_sel:
            raise ValueError(f"No matching cases in {wanted}. Available: {[c[0] for c in CASES]}")
    else:
        cases_sel = CASES
    all_rows = [run_case(name, maker, sig, outdir=args.outdir, trials=args.trials, seed=args.seed, solvers=solvers,
                         max_iter=args.max_iter, S_window=args.S_window, use_central_diff=bool(args.use_central_diff),
                         success_tol=args.success_tol) for name, maker, sig in cases_sel]
    df_all = pd.concat(all_rows, ignore_index=True)
    df_all.to_csv(os.path.join(args.outdir, "summary_all.csv"), index=False)
    _aggregate(df_all).to_csv(os.path.join(args.outdir, "summary_agg.csv"), index=False)
    print("\nSaved:")
    print(" ", os.path.join(args.outdir, "summary_all.csv"))
    print(" ", os.path.join(args.outdir, "summary_agg.csv"))
    for name, _, _ in cases_sel:
        print(" ", os.path.join(args.outdir, name, "summary_all.csv"))


if __name__ == "__main__":
    main()
