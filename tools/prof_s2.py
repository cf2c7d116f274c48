#!/usr/bin/env python
"""Runs the LQR-boundary entry point (hop_select_f64) on a synthetic S2 batch; used under ncu / for A-B timing.
  python tools/prof_s2.py --d 4 --m 2 --N 128 --B 131072 [--reps 3]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "time-opt-ilqr_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from hop import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--d", type=int, default=4)
ap.add_argument("--m", type=int, default=2)
ap.add_argument("--N", type=int, default=128)
ap.add_argument("--B", type=int, default=131072)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
d, m, N, B = a.d, a.m, a.N, a.B
gen = torch.Generator(device=dev); gen.manual_seed(1234)
A = torch.eye(d, dtype=torch.float64, device=dev).expand(B, N, d, d) + \
    0.05 / np.sqrt(d) * torch.randn((B, N, d, d), dtype=torch.float64, device=dev, generator=gen)
Bm = 0.05 * torch.randn((B, N, d, m), dtype=torch.float64, device=dev, generator=gen)
Q = torch.diag_embed(0.5 + 1.5 * torch.rand((B, N, d), dtype=torch.float64, device=dev, generator=gen))
QT = (50.0 * torch.eye(d, dtype=torch.float64, device=dev)).expand(B, N, d, d).contiguous()
Rd = 0.05 + 0.45 * torch.rand((B, m), dtype=torch.float64, device=dev, generator=gen)
Rinv = torch.diag_embed(1.0 / (Rd + 1e-9))
z0 = torch.randn((B, d), dtype=torch.float64, device=dev, generator=gen)
w = 0.01 + 0.09 * torch.rand((B,), dtype=torch.float64, device=dev, generator=gen)
run = lambda: api.propagator_all_Jt_aug_batched(A, Bm, Q, Rinv, z0, QT, 1, N, w_explicit=w, mode=api.MODE_FAST)  # noqa: E731
sel = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    sel = run()
e1.record()
torch.cuda.synchronize()
dt = e0.elapsed_time(e1) * 1e-3 / a.reps
per = 8 * (N * (3 * d * d + d * m) + m * m + d) + 8 * N + 4
print(json.dumps({"d": d, "m": m, "N": N, "B": B, "ms": dt * 1e3, "solves_per_s": B / dt, "GBs": B * per / dt / 1e9,
                  "J_sum": float(sel.J.sum()), "T_sum": int(sel.T_star.sum()), "status_nonzero": int((sel.status != 0).sum())}))
