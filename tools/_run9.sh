mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for c in Segway_Balance Cartpole_SwingUp DoubleIntegrator; do for B in 25 4096 131072; do for v in 0 1; do
  echo -n "lanes=$v "; HOP_FUSED_LANES=$v python tools/prof_small.py --case $c --B $B 2>&1 | tail -1
done; done; done
timeout 900 python tests/run_configs.py --configs 1,2,3 > gpurun_out/cfg_epl.jsonl 2> gpurun_out/cfg_epl.err; echo "cfg rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/cfg_epl.jsonl"):
    d=json.loads(l)
    if d["config"]==1: print(d); continue
    print(d["config"], "device_s %.4f"%d["device_s"], {k:round(v,4) for k,v in d["phase_seconds_rank0"].items()}, d["parity_vs_oracle"])
PY
