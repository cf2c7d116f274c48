#!/bin/bash
# Round-2 GPU pass E: line-search variants (0 = split single-warp CTAs, 2 = six roles per CTA) on configs 2, 3 and on config 4
# with 2048 instances (the per-GPU load of the 8-GPU strong-scaling run), then the GPU tests.
mkdir -p gpurun_out
for v in 0 2; do
  HOP_LS_VARIANT=$v timeout 900 python tests/run_configs.py --configs 2,3 > gpurun_out/r2e_cfg23_ls$v.jsonl 2> gpurun_out/r2e_cfg23_ls$v.err
  HOP_LS_VARIANT=$v HOP_CFG4_B=2048 timeout 900 python tests/run_configs.py --configs 4 > gpurun_out/r2e_cfg4_2048_ls$v.jsonl 2> gpurun_out/r2e_cfg4_2048_ls$v.err
  HOP_LS_VARIANT=$v HOP_CFG4_B=8192 timeout 900 python tests/run_configs.py --configs 4 > gpurun_out/r2e_cfg4_8192_ls$v.jsonl 2> gpurun_out/r2e_cfg4_8192_ls$v.err
  python - <<PY
import json
for f in ("gpurun_out/r2e_cfg23_ls$v.jsonl","gpurun_out/r2e_cfg4_2048_ls$v.jsonl","gpurun_out/r2e_cfg4_8192_ls$v.jsonl"):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print("LS variant $v config", d["config"], d["instances"], "%.4f s"%d["device_s"], {k:round(x*1e3,2) for k,x in d["phase_seconds_rank0"].items()}, d["parity_vs_oracle"]["fast"]["T_hist_identical"], d["parity_vs_oracle"]["exact"]["T_hist_identical"])
PY
done
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2e_pytest.log | cut -c1-300
