#!/bin/bash
mkdir -p gpurun_out
for v in 0 2; do
  HOP_BW_VARIANT=$v timeout 900 python tests/run_configs.py --configs 4 > gpurun_out/q3_cfg4_bw$v.jsonl 2> gpurun_out/q3_cfg4_bw$v.err
  HOP_BW_VARIANT=$v HOP_CFG4_B=2048 timeout 900 python tests/run_configs.py --configs 4 > gpurun_out/q3_cfg4_2048_bw$v.jsonl 2>> gpurun_out/q3_cfg4_bw$v.err
  python - <<PY
import json
for f in ("gpurun_out/q3_cfg4_bw$v.jsonl","gpurun_out/q3_cfg4_2048_bw$v.jsonl"):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print("BW variant $v config", d["config"], d["instances"], "%.4f s"%d["device_s"], {k:round(x*1e3,2) for k,x in d["phase_seconds_rank0"].items()}, d["parity_vs_oracle"]["fast"]["T_hist_identical"], d["parity_vs_oracle"]["exact"]["T_hist_identical"], d["parity_vs_oracle"]["fast"]["max_rel_J_hist_where_T_identical"])
PY
done
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/q3_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/q3_pytest.log | cut -c1-300
