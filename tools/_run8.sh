mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
HOP_LS_SERIAL=1 timeout 900 python tests/run_configs.py --configs 2,3,4 > gpurun_out/cfg_serial.jsonl 2> gpurun_out/cfg_serial.err; echo "serial rc=$?"
timeout 900 python tests/run_configs.py --configs 2,3,4 > gpurun_out/cfg_par.jsonl 2> gpurun_out/cfg_par.err; echo "par rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/cfg_serial.jsonl","gpurun_out/cfg_par.jsonl"):
    for l in open(f):
        d=json.loads(l); print(f[-14:], d["config"], "device_s %.4f"%d["device_s"], {k:round(v,4) for k,v in d["phase_seconds_rank0"].items()}, d["parity_vs_oracle"])
PY
