#!/bin/bash
# Round-2 GPU pass G: small-system configurations after the Newton reciprocal in the d <= 5 Gauss-Jordan kernels + GPU tests
mkdir -p gpurun_out
timeout 900 python tests/run_configs.py --configs 2,3 > gpurun_out/r2g_cfg23.jsonl 2> gpurun_out/r2g_cfg23.err
python - <<'PY'
import json
for l in open("gpurun_out/r2g_cfg23.jsonl"):
    if l.startswith("{"):
        d=json.loads(l); p=d["parity_vs_oracle"]; print("config", d["config"], d["instances"], "%.4f s"%d["device_s"], {k:round(x*1e3,2) for k,x in d["phase_seconds_rank0"].items()}, "well", p["well_posed"], "fast", p["fast"]["T_hist_identical"], p["fast"]["T_hist_identical_among_well_posed"], "exact", p["exact"]["T_hist_identical"], p["exact"]["T_hist_identical_among_well_posed"])
PY
for d in 3 4 5; do m=1; [ $d = 4 ] && m=2; timeout 300 python tools/prof_s2.py --d $d --m $m --N 128 --B 524288 --reps 3 2>&1 | tail -1; done
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2g_pytest.log | cut -c1-300
