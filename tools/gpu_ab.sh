#!/bin/bash
# A/B of the selection-kernel variants on one B200 (run under gpurun).  Output: gpurun_out/ab_*.log
# usage: gpu_ab.sh [--notest] name:ENV=VAL[,ENV=VAL] ...
mkdir -p gpurun_out
if [ "$1" == "--notest" ]; then shift; else
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log; fi
for v in "$@"; do
  name=${v%%:*}; envs=$(echo ${v#*:} | tr ',' ' ')
  env $envs timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/ab_$name.log 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$name.log").read().strip().splitlines()[-1])
    print("$name", "value %.0f"%d["value"], "ms %.2f"%d["ms_per_step"], "e2e %.0f"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"], d["cpu_baseline"]["parity_on_sample"], d["clocks"])
except Exception as e:
    print("$name failed", e)
PY
done
