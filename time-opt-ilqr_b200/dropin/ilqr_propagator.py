# -*- coding: utf-8 -*-
"""Legacy-name adapter for the reference's self-contained monolith ilqr_propagator.py.

The monolith duplicates the modular code with older constants.  This module exposes the
monolith's function names on top of the B200 path; where the monolith's constants differ from the modular
code the differences that the device ABI can express are honoured (forward-difference step eps=1e-6 without
the relative term, explicit Qtilde in build_terminal_aug_list, 12-tuples from make_*), the others (4-try
jitter ladder, missing finiteness guards) follow the modular semantics.  Ballbot and the plotting / main()
driver are out of scope."""
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
chol_inv = _utils.chol_inv
chol_solve = _utils.chol_solve
angle_normalize = _utils.angle_normalize
wrap_error = _utils.wrap_error
linearize_central_diff_traj = _lin.linearize_central_diff_traj
compute_affine_residuals = _lin.compute_affine_residuals
propagator_all_Jt_aug = _hs.propagator_all_Jt_aug
rollout = _solver.rollout
backward_pass_truncated = _solver.backward_pass_truncated
forward_linesearch_fixedT = _solver.forward_linesearch_fixedT


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


def ilqr_timeopt(F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, method="propagator", max_iter=20, lm_init=1e-3,
                 S_window=10, use_central_diff=True, wrap_idx=None, **kw):
    return _solver.ilqr_timeopt(F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, method=method, max_iter=max_iter,
                                lm_init=lm_init, S_window=S_window, use_central_diff=use_central_diff, wrap_idx=wrap_idx, **kw)


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
