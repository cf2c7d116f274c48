# -*- coding: utf-8 -*-
"""Drop-in for the reference's augmented.py: the homogeneous augmented-state embedding.

build_augmented_sequence_QR / build_terminal_aug_list keep the reference signatures and return lists of
ndarrays (blocks built by hop_build_augmented_f64 / hop_build_terminal_f64).  The fused solver path never
materialises these blocks (hop_select_fused_f64 builds them in registers); these functions exist for API
parity and for callers that want to inspect the blocks."""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from _bridge import _cabi, api, dev, ptr, stack, stream, torch
from linearization import compute_affine_residuals
from utils import _sym, as_terminal_weight, chol_inv


def build_augmented_sequence_QR(F, A_list, B_list, X, U, xg, u_ref, Q, R, w, wrap_idx: Optional[List[int]] = None,
                                q_reg: float = 1e-9, rho_reg: float = 1e-12, extra_stage_cost=None):
    if extra_stage_cost is not None:
        raise NotImplementedError("extra_stage_cost (Python callback) is not supported on the B200 path")
    lib = _cabi.require_device()
    X = np.asarray(X, dtype=float)
    U = np.asarray(U, dtype=float).reshape(len(A_list), -1)
    N, n, m = len(A_list), X.shape[1], U.shape[1]
    d = n + 1
    R = _sym(np.asarray(R, dtype=float))
    R_inv = chol_inv(R)                                                        # augmented.py:23
    a = stack(compute_affine_residuals(F, X, U)).reshape(1, N, n)
    At, Bt, at = dev(stack(A_list)[None]), dev(stack(B_list)[None]), dev(a)
    Xt, Ut = dev(X[None]), dev(U[None])
    xgt, wt = dev(np.asarray(xg, dtype=float).reshape(1, n)), dev([float(w)])
    urt, Qt_ = dev(u_ref), dev(Q)        # keep every device buffer referenced until the launch has been issued
    A_aug = torch.empty((1, N, d, d), dtype=torch.float64, device=At.device)
    B_aug = torch.empty((1, N, d, m), dtype=torch.float64, device=At.device)
    Q_aug = torch.empty((1, N, d, d), dtype=torch.float64, device=At.device)
    _cabi.check(lib.hop_build_augmented_f64(1, N, n, m, ptr(At), ptr(Bt), ptr(at), ptr(Xt), ptr(Ut), N * m, ptr(xgt), ptr(wt),
                                            ptr(urt), ptr(Qt_), api.wrap_mask(wrap_idx), float(q_reg),
                                            float(rho_reg), ptr(A_aug), ptr(B_aug), ptr(Q_aug), stream()),
                "hop_build_augmented_f64")
    A_aug, B_aug, Q_aug = A_aug[0].cpu().numpy(), B_aug[0].cpu().numpy(), Q_aug[0].cpu().numpy()
    z0 = np.zeros(d)
    z0[-1] = 1.0
    return ([A_aug[k] for k in range(N)], [B_aug[k] for k in range(N)], [Q_aug[k] for k in range(N)], [R] * N, z0, R_inv)


def build_terminal_aug_list(X, xg, alpha, wrap_idx: Optional[List[int]] = None, rho_reg: float = 1e-12):
    lib = _cabi.require_device()
    X = np.asarray(X, dtype=float)
    n = X.shape[1]
    N = X.shape[0] - 1
    Qf = as_terminal_weight(alpha, n)
    Xt = dev(X[None])
    xgt, Qft = dev(np.asarray(xg, dtype=float).reshape(1, n)), dev(Qf)
    QT = torch.empty((1, N, n + 1, n + 1), dtype=torch.float64, device=Xt.device)
    _cabi.check(lib.hop_build_terminal_f64(1, N, n, ptr(Xt), ptr(xgt), ptr(Qft),
                                           api.wrap_mask(wrap_idx), float(rho_reg), ptr(QT), stream()),
                "hop_build_terminal_f64")
    QT = QT[0].cpu().numpy()
    return [QT[t] for t in range(N)]
