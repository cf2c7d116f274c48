# -*- coding: utf-8 -*-
"""Legacy-name adapter for the reference's self-contained monolith ilqr_propagator.py.

The monolith duplicates the modular code with older constants.  This module exposes the
monolith's function names on top of the B200 path; where the monolith's constants differ from the modular
code the differences that the device ABI can express are honoured (forward-difference step eps=1e-6 without
the relative term, explicit Qtilde in build_terminal_aug_list, 12-tuples from make_*, the 4-try ladder of the
stand-alone chol_inv / chol_solve, the post-solve consistency check), the others (np.linalg.lstsq fallback,
missing finiteness guards) follow the modular semantics and say so.  The values the monolith shipped
(plots/summary.csv) are reproduced by tests/test_dropin_gpu.py.  Ballbot and the plotting / main() driver are
out of scope."""
from __future__ import annotations

import numpy as np

import augmented as _aug
import horizon_selection as _hs
import linearization as _lin
import solver as _solver
import systems as _sys
import utils as _utils
from _bridge import cases

_sym = _utils._sym
angle_normalize = _utils.angle_normalize
wrap_error = _utils.wrap_error
linearize_central_diff_traj = _lin.linearize_central_diff_traj
compute_affine_residuals = _lin.compute_affine_residuals
propagator_all_Jt_aug = _hs.propagator_all_Jt_aug
rollout = _solver.rollout
backward_pass_truncated = _solver.backward_pass_truncated
forward_linesearch_fixedT = _solver.forward_linesearch_fixedT
bruteforce_all_Jt_backward_expansion = _solver.bruteforce_all_Jt_backward_expansion


def chol_inv(A, jitter: float = 1e-9, max_tries: int = 4):
    """ilqr_propagator.py:21-31: FOUR Cholesky attempts (jitter x10 each), then the plain inverse of A + eps I
    (hop_chol_inv_f64 takes the number of tries; its fallback is the LU inverse of A + eps I, i.e. np.linalg.inv).
    Deviation: a non-finite input raises FloatingPointError (the monolith has no finiteness guard)."""
    return _utils.chol_inv(A, jitter=jitter, max_tries=max_tries)


def chol_solve(A, B, jitter: float = 1e-9, max_tries: int = 4):
    """ilqr_propagator.py:33-43: four attempts.  Deviation: where the monolith falls back to np.linalg.lstsq after the
    ladder, this raises LinAlgError as the modular utils.chol_solve does (there is no host linear algebra on this path)."""
    return _utils.chol_solve(A, B, jitter=jitter, max_tries=max_tries)


def linearize_forward_diff_traj(F, X, U, epsx=1e-6, epsu=1e-6):
    """ilqr_propagator.py:142-152: constant step (no relative term)."""
    return _lin.linearize_forward_diff_traj(F, X, U, epsx=epsx, epsu=epsu, relx=0.0, relu=0.0)


def build_augmented_sequence_QR(F, A_list, B_list, X, U, xg, u_ref, Q, R, w, wrap_idx=None, q_reg=1e-9, rho_reg=1e-12):
    return _aug.build_augmented_sequence_QR(F, A_list, B_list, X, U, xg, u_ref, Q, R, w, wrap_idx, q_reg, rho_reg)


def build_terminal_aug_list(X, xg, alpha, Qtilde, wrap_idx=None, rho_reg=1e-12):
    """ilqr_propagator.py:194-207: terminal weight P = sym(alpha * Qtilde)."""
    return _aug.build_terminal_aug_list(X, xg, _sym(np.asarray(alpha, float) * np.asarray(Qtilde, float)), wrap_idx, rho_reg)


def cost_timeopt_true(X, U, xg, u_ref, Q, R, alpha, w, T_star, wrap_idx=None):
    return _solver.cost_timeopt_true(X, U, xg, u_ref, Q, R, alpha, w, T_star, wrap_idx)


def ilqr_timeopt(F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, method: str = "propagator", max_iter: int = 15,
                 lm_init: float = 1e-3, S_window: int = 20, S_left=None, S_right=None, wrap_idx=None,
                 use_central_diff: bool = True):
    """ilqr_propagator.py:459-658 (method="propagator"): the same outer loop as the modular solver (warm start, accept /
    reject, LM schedule, stop rule), then the monolith's own post-solve check: re-linearise the final trajectory and compare
    the propagator curve with the brute-force curve (ilqr_propagator.py:630-643) -> result["consistency_check"].
    Inside the solve the device kernels use the modular 8-try ladder (the monolith's 4 tries + np.linalg.inv only differ
    when a Cholesky fails five times; its shipped cases never climb the ladder)."""
    res = _solver.ilqr_timeopt(F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, method=method, max_iter=max_iter,
                               lm_init=lm_init, S_window=S_window, use_central_diff=use_central_diff, wrap_idx=wrap_idx)
    X, U = res["X"], res["U"]
    A_list, B_list = (linearize_central_diff_traj if use_central_diff else linearize_forward_diff_traj)(F, X, U)
    J_back = bruteforce_all_Jt_backward_expansion(A_list, B_list, X, U, xg, u_ref, Q, R, alpha, w, T_max, wrap_idx=wrap_idx)
    A_aug, B_aug, Q_aug, R_list, z0, R_inv = build_augmented_sequence_QR(F, A_list, B_list, X, U, xg, u_ref, Q, R, w, wrap_idx=wrap_idx)
    QT_list = build_terminal_aug_list(X, xg, alpha, np.eye(X.shape[1]), wrap_idx=wrap_idx)
    J_prop = propagator_all_Jt_aug(A_aug, B_aug, Q_aug, R_list, z0, QT_list, T_use=T_max, R_inv_cached=R_inv)
    diff = J_prop - J_back
    out = {k: res[k] for k in ("X", "U", "J_hist", "T_hist", "timers", "T_star")}
    out["J_curve"] = J_prop
    out["consistency_check"] = {"max_abs_diff": float(np.max(np.abs(diff))), "rmse": float(np.sqrt(np.mean(diff ** 2)))}
    return out


def _twelve(t):
    return t[:12]   # the monolith's make_* return 12-tuples ending with wrap_idx (ilqr_propagator.py:668)


def make_double_integrator(dt=0.05, N=120):
    return _twelve(_sys.make_double_integrator(dt, N))


def make_quadrotor(dt=0.05, N=160):
    return _twelve(_sys.make_quadrotor(dt, N))


def make_segway(dt=0.02, N=240):
    """Legacy Segway constants (ilqr_propagator.py:670-683): x0=[2,0,2,0], scalar alpha=120."""
    t = list(_sys.make_segway_balance(dt, N))
    t[1] = np.array([2.0, 0.0, 2.0, 0.0])
    t[6] = 120.0
    return _twelve(tuple(t))


def make_ballbot(*_a, **_k):
    raise NotImplementedError("Ballbot is not part of the reference's CASES and has no device dynamics")
