# -*- coding: utf-8 -*-
"""Drop-in for the reference's horizon_selection.py: the information-form propagator sweep (HOP).

propagator_all_Jt_aug keeps the reference signature (horizon_selection.py:36-86) and returns J[T_use];
`propagator_all_Jt_aug_batched` is the additive batched entry point.  The one-pass helpers
(value_expansions_and_gains_prefix, onepass_pick_T_singlepass) belong to baseline2 and are out of scope."""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from _bridge import api, dev, raise_status, stack
from utils import chol_inv

propagator_all_Jt_aug_batched = api.propagator_all_Jt_aug_batched


def propagator_all_Jt_aug(A_aug: List[np.ndarray], B_aug: List[np.ndarray], Q_aug: List[np.ndarray], R_list: List[np.ndarray],
                          z0: np.ndarray, QT_aug_list: List[np.ndarray], T_use: Optional[int] = None,
                          R_inv_cached: Optional[np.ndarray] = None) -> np.ndarray:
    """Compute J(T) for all T using the information-form propagator (on the GPU)."""
    N = len(A_aug) if T_use is None else int(T_use)
    if N <= 0:
        return np.zeros(0, dtype=float)
    if R_inv_cached is not None:
        R_inv = np.asarray(R_inv_cached, dtype=float)
    elif all(R_list[k] is R_list[0] for k in range(N)):
        R_inv = chol_inv(R_list[0])
    else:
        R_inv = np.stack([chol_inv(R_list[k]) for k in range(N)])[None]         # [1, N, m, m]
    sel = api.propagator_all_Jt_aug_batched(dev(stack(A_aug[:N])[None]), dev(stack(B_aug[:N])[None]), dev(stack(Q_aug[:N])[None]),
                                            dev(R_inv), dev(np.asarray(z0, dtype=float).reshape(-1)),
                                            dev(stack(QT_aug_list[:N])[None]), 1, N)
    raise_status(int(sel.status[0]), "chol_inv(A)")
    return sel.J[0].cpu().numpy()


def _baseline2(*_a, **_k):
    raise NotImplementedError("one-pass horizon selection (baseline2) is a competitor method, outside the HOP hot path")


value_expansions_and_gains_prefix = _baseline2
onepass_pick_T_singlepass = _baseline2
