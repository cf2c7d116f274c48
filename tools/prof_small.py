#!/usr/bin/env python
"""Times the x0 -> T* selection pipeline of a small system (fused selection kernel) at a given batch; A/B tool.
  python tools/prof_small.py --case Cartpole_SwingUp --B 4096 [--reps 5]"""
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
ap.add_argument("--case", default="Cartpole_SwingUp")
ap.add_argument("--B", type=int, default=4096)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
case = cases.make_case(a.case)
x0 = case[1]
x0s = torch.as_tensor(x0[None] + 0.1 * np.random.default_rng(0).standard_normal((a.B, x0.size)), device="cuda:0")
sel = api.HorizonSelector(case, a.B, device=x0s.device, mode=api.MODE_EXACT)
r = sel(x0s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    r = sel(x0s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
print(json.dumps({"case": a.case, "B": a.B, "ms_pipeline": ms, "solves_per_s": a.B / ms * 1e3, "T_sum": int(r.T_star.sum()),
                  "J_sum": float(torch.nan_to_num(r.J).sum())}))
