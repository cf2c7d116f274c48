#!/usr/bin/env python
"""A/B tool: times the public host-buffer entry (hop.api.select_horizon_host, 65 536 quadrotor instances from pinned host
buffers, outputs to pinned host buffers) with wall-clock time around synchronous calls; environment switches such as
HOP_HOST_CHUNKS / HOP_LIN_MINB select the variant.  Not a bench value (bench.py reports e2e)."""
import os, sys, json, time
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "time-opt-ilqr_b200"))
import numpy as np, torch
from hop import api, cases
case = cases.make_case("Quadrotor", N=128)
B=65536
x0 = case[1]
sigma = np.array([0.4,0.4,0.4]+[0.0]*9)
x0p = torch.empty((B,12), dtype=torch.float64, pin_memory=True)
x0p.copy_(torch.as_tensor(x0[None] + sigma[None]*np.random.default_rng(0).standard_normal((B,12))))
T_max=128
out = (torch.empty((B,T_max), dtype=torch.float64, pin_memory=True).numpy(), torch.empty(B, dtype=torch.int32, pin_memory=True).numpy(),
       torch.empty(B, dtype=torch.float64, pin_memory=True).numpy(), torch.empty(B, dtype=torch.int32, pin_memory=True).numpy())
for _ in range(3): api.select_horizon_host(case, x0p.numpy(), mode=api.MODE_FAST, out=out)
t=time.perf_counter()
K=6
for _ in range(K): api.select_horizon_host(case, x0p.numpy(), mode=api.MODE_FAST, out=out)
dt=(time.perf_counter()-t)/K
print(json.dumps({"chunks": os.environ.get("HOP_HOST_CHUNKS"), "ms": dt*1e3, "e2e": B/dt, "T_sum": int(out[1].sum())}))
