#!/usr/bin/env python
"""Runs a short batched HOP-DDP solve of a small system (for ncu captures of the line-search / backward kernels).
  python tools/prof_ddp.py --case Cartpole_SwingUp --B 4096 --iters 3"""
import argparse
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
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
case = cases.make_case(a.case)
x0 = case[1]
x0s = torch.as_tensor(x0[None] + 0.1 * np.random.default_rng(0).standard_normal((a.B, x0.size)), device="cuda:0")
out = api.ilqr_timeopt_batched(case, x0s, max_iter=a.iters, use_central_diff=False, mode=api.MODE_EXACT)
torch.cuda.synchronize()
print("T* in", int(out["T_star"].min()), int(out["T_star"].max()))
