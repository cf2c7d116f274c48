#!/usr/bin/env python
"""Runs the headline S1 selection (quadrotor, N=128) from device-resident x0; used under ncu / for A-B timing.
  python tools/prof_s1.py --B 8192 [--reps 3] [--mode fast|exact]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "time-opt-ilqr_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from hop import api, cases  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=8192)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--mode", default="fast")
a = ap.parse_args()
case = cases.make_case("Quadrotor", N=128)
x0 = case[1]
sigma = np.array([0.4, 0.4, 0.4] + [0.0] * 9)
x0s = torch.as_tensor(x0[None] + sigma[None] * np.random.default_rng(0).standard_normal((a.B, 12)), device="cuda:0")
sel = api.HorizonSelector(case, a.B, device=x0s.device, mode=api.MODE_FAST if a.mode == "fast" else api.MODE_EXACT)
r = sel(x0s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    r = sel(x0s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
print(json.dumps({"B": a.B, "ms_pipeline": ms, "solves_per_s": a.B / ms * 1e3, "T_sum": int(r.T_star.sum()), "J_sum": float(r.J.sum())}))
