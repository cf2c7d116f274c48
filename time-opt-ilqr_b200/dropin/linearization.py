# -*- coding: utf-8 -*-
"""Drop-in for the reference's linearization.py (finite-difference Jacobians + affine residuals).

linearize_{central,forward}_diff_traj and compute_affine_residuals keep the reference signatures and
return Python lists of ndarrays; the work runs on the GPU (hop_linearize_f64 / hop_affine_residuals_f64).
The negative-time prefix helpers (extend_nominal_backward, ...) only serve the one-pass baseline, which
is out of scope for the HOP path."""
from __future__ import annotations

import ctypes as C

import numpy as np

from _bridge import _cabi, api, dev, ptr, require_dynamics, stream, torch


def _lin(F, X, U, central, epsx, epsu, relx, relu):
    require_dynamics(F)
    X = np.asarray(X, dtype=float)
    U = np.asarray(U, dtype=float).reshape(len(U), -1)
    A, B = api.linearize_batched(F, dev(X[None]), dev(U[None]), central=central, epsx=epsx, epsu=epsu, relx=relx, relu=relu)
    A, B = A[0].cpu().numpy(), B[0].cpu().numpy()
    return [A[k] for k in range(A.shape[0])], [B[k] for k in range(B.shape[0])]


def linearize_central_diff_traj(F, X, U, epsx: float = 1e-5, epsu: float = 1e-5, relx: float = 1e-6, relu: float = 1e-6):
    return _lin(F, X, U, True, epsx, epsu, relx, relu)


def linearize_forward_diff_traj(F, X, U, epsx: float = 1e-5, epsu: float = 1e-5, relx: float = 1e-6, relu: float = 1e-6):
    return _lin(F, X, U, False, epsx, epsu, relx, relu)


def compute_affine_residuals(F, X: np.ndarray, U: np.ndarray):
    require_dynamics(F)
    lib = _cabi.require_device()
    X = np.asarray(X, dtype=float)
    U = np.asarray(U, dtype=float).reshape(len(U), -1)
    N, n, m = U.shape[0], X.shape[1], U.shape[1]
    Xt, Ut = dev(X[None]), dev(U[None])
    a = torch.empty((1, N, n), dtype=torch.float64, device=Xt.device)
    p = api._params(F)
    _cabi.check(lib.hop_affine_residuals_f64(1, F.hop_sys, p.ctypes.data_as(C.c_void_p), N, ptr(Xt), ptr(Ut), N * m, ptr(a),
                                             stream()), "hop_affine_residuals_f64")
    a = a[0].cpu().numpy()
    return [a[k].reshape(-1, 1) for k in range(N)]


def extend_nominal_backward(*_a, **_k):
    raise NotImplementedError("negative-time prefix extension only serves the one-pass baseline (baseline2); "
                              "it is outside the HOP hot path implemented on the B200")
