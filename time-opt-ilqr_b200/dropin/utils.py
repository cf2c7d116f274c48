# -*- coding: utf-8 -*-
"""Drop-in for the reference's utils.py (linear algebra + angle wrapping) on the B200 path.

Same names, signatures, defaults and exception types as utils.py:35-137.  `chol_inv` / `chol_solve`
run on the GPU (hop_chol_inv_f64 / hop_chol_solve_f64: Cholesky route, 1e-9 jitter ladder x10 up to
8 tries, LU fallback for the inverse only); the scalar helpers are plain host arithmetic.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from _bridge import _cabi, api, dev, ptr, raise_status, stream, torch


def _sym(A: np.ndarray) -> np.ndarray:
    return 0.5 * (A + A.T)


def _assert_finite(name: str, X: np.ndarray):
    if not np.all(np.isfinite(X)):
        raise FloatingPointError(f"Non-finite values in {name}")


def as_terminal_weight(alpha, n: int) -> np.ndarray:
    return api.as_terminal_weight(alpha, n)


def chol_inv(A: np.ndarray, jitter: float = 1e-9, max_tries: int = 8) -> np.ndarray:
    lib = _cabi.require_device()
    A = np.asarray(A, dtype=float)
    d = A.shape[0]
    At = dev(A.reshape(1, d, d))
    X = torch.empty_like(At)
    st = torch.zeros(1, dtype=torch.int32, device=At.device)
    _cabi.check(lib.hop_chol_inv_f64(1, d, ptr(At), ptr(X), float(jitter), int(max_tries), ptr(st), stream()), "hop_chol_inv_f64")
    raise_status(int(st[0]), "chol_inv(A)")
    return X[0].cpu().numpy()


def chol_solve(A: np.ndarray, B: np.ndarray, jitter: float = 1e-9, max_tries: int = 8) -> np.ndarray:
    lib = _cabi.require_device()
    A = np.asarray(A, dtype=float)
    B0 = np.asarray(B, dtype=float)
    d = A.shape[0]
    Bm = B0.reshape(d, -1)
    At, Bt = dev(A.reshape(1, d, d)), dev(Bm.reshape(1, d, -1))
    X = torch.empty_like(Bt)
    st = torch.zeros(1, dtype=torch.int32, device=At.device)
    _cabi.check(lib.hop_chol_solve_f64(1, d, Bm.shape[1], ptr(At), ptr(Bt), ptr(X), float(jitter), int(max_tries), ptr(st),
                                       stream()), "hop_chol_solve_f64")
    code = int(st[0]) & 0xFF
    if code == 1:
        raise FloatingPointError("Non-finite values in chol_solve")
    if code == 2:
        raise np.linalg.LinAlgError(f"chol_solve failed: matrix not PD after jitter up to {jitter * 10 ** max_tries:g}")
    return X[0].cpu().numpy().reshape(B0.shape)


def angle_normalize(a: float) -> float:
    return (a + np.pi) % (2.0 * np.pi) - np.pi


def wrap_error(e: np.ndarray, wrap_idx: Optional[List[int]] = None) -> np.ndarray:
    if not wrap_idx:
        return e
    e = np.asarray(e, dtype=float).copy()
    for i in wrap_idx:
        e[i] = angle_normalize(float(e[i]))
    return e
