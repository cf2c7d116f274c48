#!/usr/bin/env python
"""Latency of one LQR-boundary horizon selection (hop_select_f64) for SMALL batches: sequential sweep vs HOP_MODE_SCAN."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "time-opt-ilqr_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle as O
from _common import s2_batch
from hop import api
d, m, N = 13, 4, 128
for B in (1, 8, 148, 1184, 16384):
    A, Bm, Q, R, z0, w, QT = s2_batch(range(min(B, 16)), d, m, N)
    rep = -(-B // A.shape[0])
    t = lambda x: torch.as_tensor(np.concatenate([x] * rep)[:B], device="cuda")
    Rinv = np.stack([O.chol_inv(r) for r in R])
    args = (t(A), t(Bm), t(Q), t(Rinv), t(z0), t(QT), 1, N)
    out = {}
    for name, mode in (("sequential", api.MODE_EXACT), ("scan", api.MODE_SCAN)):
        for _ in range(3): api.propagator_all_Jt_aug_batched(*args, mode=mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): s = api.propagator_all_Jt_aug_batched(*args, mode=mode)
        e1.record(); torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / 10
    print(json.dumps({"d": d, "N": N, "batch": B, "ms_sequential": out["sequential"], "ms_scan": out["scan"],
                      "speedup": out["sequential"] / out["scan"]}))
