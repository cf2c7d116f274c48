#!/bin/bash
for e in 1024 8192; do
HOP_EPL_MAX_BATCH=$e python tests/run_configs.py --configs 3 2>/dev/null > gpurun_out/tmp_cfg3_$e.jsonl
python - <<PY
import json
for l in open("gpurun_out/tmp_cfg3_$e.jsonl"):
    if l.startswith("{"):
        d=json.loads(l); print("EPL_MAX_BATCH=$e: config", d["config"], "%.4f s"%d["device_s"], {k:round(x*1e3,2) for k,x in d["phase_seconds_rank0"].items()})
PY
done
