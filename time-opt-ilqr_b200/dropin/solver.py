# -*- coding: utf-8 -*-
"""Drop-in for the reference's solver.py (time-optimal iLQR with horizon selection) on the B200 path.

Same public names and signatures as solver.py:42-105,156-286,449-779.  Only method="propagator" (HOP,
"ourmethod") runs here; "bruteforce" (baseline1) and "onepass" (baseline2) are competitor methods outside
the hot path and raise NotImplementedError.  `ilqr_timeopt_batched` is the additive batched entry point:
the whole per-instance state machine (warm start, accept/reject, LM schedule, stop rule) runs on the
device for a batch of (x0, xg, w)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from _bridge import api, cases, dev, raise_status, require_dynamics, stack, torch
from utils import as_terminal_weight

ilqr_timeopt_batched = api.ilqr_timeopt_batched


def _case_tuple(F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx):
    return (require_dynamics(F), np.asarray(x0, float), np.asarray(xg, float), np.asarray(u_ref, float), np.asarray(Q, float),
            np.asarray(R, float), alpha, float(w), int(N), int(T_min), int(T_max), list(wrap_idx or []), None)


def rollout(F, x0: np.ndarray, U: np.ndarray, *, max_state_norm: float = 1e6) -> np.ndarray:
    """Roll forward dynamics with simple divergence checks (solver.py:42-62)."""
    require_dynamics(F)
    x0 = np.asarray(x0, dtype=float).reshape(1, -1)
    U = np.asarray(U, dtype=float)
    if U.ndim == 1:
        U = U.reshape(-1, 1)
    return api.rollout_batched(F, dev(x0), dev(U[None]), max_state_norm)[0].cpu().numpy()


def _cost_case(X, U, xg, u_ref, Q, R, alpha, w):
    n, m = X.shape[1], U.shape[1]
    sys_id = {(2, 1): 0, (4, 1): 1, (12, 4): 2}.get((n, m))      # the cost / backward kernels only need (n, m)
    if sys_id is None:
        raise NotImplementedError(f"(n, m) = ({n}, {m}) is not instantiated on the device; supported: (2,1) (4,1) (12,4)")
    F = cases.Dynamics(sys_id, [0.0], 0.0, "dims-only")
    return (F, np.zeros(n), xg, u_ref, Q, R, alpha, w, U.shape[0], 1, U.shape[0], None, None)


def cost_timeopt_true(X, U, xg, u_ref, Q, R, alpha, w: float, T_star: int, wrap_idx: Optional[List[int]] = None,
                      extra_stage_cost=None) -> float:
    """True objective: running cost up to T_star + terminal cost at T_star (solver.py:65-105)."""
    if extra_stage_cost is not None:
        raise NotImplementedError("extra_stage_cost is not supported on the B200 path")
    X, U = np.asarray(X, float), np.asarray(U, float).reshape(len(U), -1)
    case = list(_cost_case(X, U, xg, u_ref, Q, R, alpha, w))
    case[11] = list(wrap_idx or [])
    J = api.cost_timeopt_true_batched(tuple(case), dev(X[None]), dev(U[None]), torch.tensor([int(T_star)], dtype=torch.int32))
    return float(J[0])


def backward_pass_truncated(A_list, B_list, X, U, xg, u_ref, Q, R, alpha, T_star: int, *, lm_lambda: float = 1e-3,
                            wrap_idx: Optional[List[int]] = None, extra_stage_cost=None):
    """Standard iLQR backward pass on [0..T_star] (solver.py:156-230): (k_list, K_list, ok)."""
    k, K, ok, _ = _backward_linesearch(None, A_list, B_list, X, U, xg, u_ref, Q, R, alpha, 0.0, T_star, lm_lambda, wrap_idx,
                                       extra_stage_cost)
    return (k, K, True) if ok else (None, None, False)


def _backward_linesearch(F, A_list, B_list, X, U, xg, u_ref, Q, R, alpha, w, T_star, lm, wrap_idx, extra_stage_cost):
    if extra_stage_cost is not None:
        raise NotImplementedError("extra_stage_cost is not supported on the B200 path")
    X, U = np.asarray(X, float), np.asarray(U, float).reshape(len(U), -1)
    T_star = int(T_star)
    if T_star <= 0:
        return None, None, False, None
    case = list(_cost_case(X, U, xg, u_ref, Q, R, alpha, w))
    if F is not None:
        case[0] = require_dynamics(F)
    case[11] = list(wrap_idx or [])
    r = api.backward_linesearch_batched(tuple(case), dev(stack(A_list)[None]), dev(stack(B_list)[None]), dev(X[None]), dev(U[None]),
                                        torch.tensor([T_star], dtype=torch.int32), float(lm))
    err = int(r["err"][0])
    if err == 1:
        raise FloatingPointError("Non-finite values in chol_solve")
    if err == 2:
        raise np.linalg.LinAlgError("chol_solve failed: matrix not PD after jitter")
    if not int(r["ok"][0]):
        return None, None, False, r
    k = r["k"][0, :T_star].cpu().numpy()
    K = r["K"][0, :T_star].cpu().numpy()
    return [k[i] for i in range(T_star)], [K[i] for i in range(T_star)], True, r


def forward_linesearch_fixedT(F, X, U, xg, u_ref, Q, R, alpha, w: float, T_star: int, k_list, K_list, *,
                              alphas: Tuple[float, ...] = (1.0, 0.5, 0.25, 0.1, 0.05), wrap_idx: Optional[List[int]] = None,
                              extra_stage_cost=None):
    """Forward pass with line-search at a fixed horizon (solver.py:233-286): (X_new, U_new, J, accepted).
    The gains are recomputed on the device together with the line search only when called through
    ilqr_timeopt; this entry point takes the caller's gains."""
    if tuple(alphas) != (1.0, 0.5, 0.25, 0.1, 0.05):
        raise NotImplementedError("the device line search uses the reference's alpha schedule (1, .5, .25, .1, .05)")
    if extra_stage_cost is not None:
        raise NotImplementedError("extra_stage_cost is not supported on the B200 path")
    import ctypes as C
    from _bridge import _cabi, ptr, stream
    require_dynamics(F)
    lib = _cabi.require_device()
    X, U = np.asarray(X, float), np.asarray(U, float).reshape(len(U), -1)
    N, n, m = U.shape[0], X.shape[1], U.shape[1]
    T_star = int(T_star)
    kl = np.zeros((1, N, m)); Kl = np.zeros((1, N, m, n))
    kl[0, :T_star] = stack(k_list).reshape(-1, m)[:T_star]
    Kl[0, :T_star] = stack(K_list).reshape(-1, m, n)[:T_star]
    d = dev
    Xt, Ut = d(X[None]), d(U[None])
    Xn, Un = torch.empty_like(Xt), torch.empty_like(Ut)
    Jn = torch.zeros(1, dtype=torch.float64, device=Xt.device); acc = torch.zeros(1, dtype=torch.int32, device=Xt.device)
    T = torch.tensor([T_star], dtype=torch.int32, device=Xt.device)
    ok = torch.ones(1, dtype=torch.int32, device=Xt.device)
    # every device buffer stays referenced until the launch has been issued
    xgt, wt, urt, Qt_, Rt, Qft = d(np.asarray(xg, float).reshape(1, n)), d([float(w)]), d(u_ref), d(Q), d(R), d(as_terminal_weight(alpha, n))
    klt, Klt = d(kl), d(Kl)
    prm = api._params(F)
    _cabi.check(lib.hop_linesearch_f64(1, F.hop_sys, prm.ctypes.data_as(C.c_void_p), N, ptr(Xt), ptr(Ut), ptr(xgt), ptr(wt), ptr(urt),
                                       ptr(Qt_), ptr(Rt), ptr(Qft), api.wrap_mask(wrap_idx), ptr(T), ptr(klt), ptr(Klt), ptr(ok),
                                       ptr(Xn), ptr(Un), ptr(Jn), ptr(acc), stream()), "hop_linesearch_f64")
    if not int(acc[0]):
        return X, U, float(Jn[0]), False
    return Xn[0].cpu().numpy(), Un[0].cpu().numpy(), float(Jn[0]), True


def bruteforce_all_Jt_backward_expansion(A_list, B_list, X, U, xg, u_ref, Q, R, alpha, w: float, T_max: int, *,
                                         lm_lambda: float = 1e-6, wrap_idx: Optional[List[int]] = None,
                                         extra_stage_cost=None) -> np.ndarray:
    """Exact J(T) curve under the iLQR quadratic model (solver.py:293-358), one warp per horizon on the device.
    The baseline-1 *solver* (method="bruteforce") stays out of scope; this curve is kept as an independent check."""
    if extra_stage_cost is not None:
        raise NotImplementedError("extra_stage_cost is not supported on the B200 path")
    X, U = np.asarray(X, float), np.asarray(U, float).reshape(len(U), -1)
    case = list(_cost_case(X, U, xg, u_ref, Q, R, alpha, w))
    case[11] = list(wrap_idx or [])
    N = len(A_list)
    J, st = api.bruteforce_all_Jt_batched(tuple(case), dev(stack(A_list)[None]), dev(stack(B_list)[None]), dev(X[None, :N + 1]),
                                          dev(U[None, :N]), T_max=int(T_max), lm_lambda=lm_lambda)
    raise_status(int(st[0]), "chol_solve(A)")
    return J[0].cpu().numpy()


def ilqr_timeopt(F, x0, xg, u_ref, Q, R, alpha, w: float, N: int, T_min: int, T_max: int, *, U_init=None,
                 method: str = "propagator", max_iter: int = 15, lm_init: float = 1e-3, S_window: int = 20,
                 wrap_idx: Optional[List[int]] = None, use_central_diff: bool = True, extra_stage_cost=None,
                 onepass_preimage: str = "fixedpoint") -> Dict[str, Any]:
    """Solve the time-penalised horizon-selection iLQR problem (solver.py:449-765) for ONE instance on the GPU."""
    assert method in ("propagator", "bruteforce", "onepass")
    if method != "propagator":
        raise NotImplementedError(f"method={method!r} is a baseline outside the HOP hot path; only 'propagator' runs on the B200")
    if extra_stage_cost is not None:
        raise NotImplementedError("extra_stage_cost is not supported on the B200 path")
    case = _case_tuple(F, x0, xg, u_ref, Q, R, alpha, w, N, T_min, T_max, wrap_idx)
    Ui = None
    if U_init is not None:                                                   # solver.py:483-490
        Ui = np.asarray(U_init, dtype=float)
        if Ui.ndim == 1:
            Ui = Ui.reshape(-1, 1)
        if Ui.shape[0] < N:
            Ui = np.vstack([Ui, np.tile(Ui[-1:], (N - Ui.shape[0], 1))])
        elif Ui.shape[0] > N:
            Ui = Ui[:N]
    if int(T_max) > int(N):
        raise IndexError("list index out of range")                           # what horizon_selection.py:59 does in the reference
    r = api.ilqr_timeopt_batched(case, dev(np.asarray(x0, float).reshape(1, -1)), U_init=None if Ui is None else dev(Ui),
                                 max_iter=max_iter, lm_init=lm_init, use_central_diff=use_central_diff)
    raise_status(int(r["status"][0]), "chol_inv(A)")
    nh = int(r["n_hist"][0])
    return {
        "X": r["X"][0].cpu().numpy(),
        "U": r["U"][0].cpu().numpy(),
        "J_hist": [float(v) for v in r["J_hist"][0, :nh].cpu().numpy()],
        "T_hist": [int(v) for v in r["T_hist"][0, :nh].cpu().numpy()],
        "timers": r["timers"],
        "J_curve": r["J_curve"][0].cpu().numpy(),
        "T_star": int(r["T_star"][0]),
        "onepass_error": None,
    }


def ilqr_timeopt_ourmethod(*args, **kwargs):
    return ilqr_timeopt(*args, method="propagator", **kwargs)


def ilqr_timeopt_baseline1(*args, **kwargs):
    return ilqr_timeopt(*args, method="bruteforce", **kwargs)


def ilqr_timeopt_baseline2(*args, **kwargs):
    return ilqr_timeopt(*args, method="onepass", **kwargs)
