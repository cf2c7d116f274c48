mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin_gpu.py -x -q -k "host or s1_from or chunked" > gpurun_out/pytest_host.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_host.log
timeout 600 python bench.py > gpurun_out/bench_chunked.json 2> gpurun_out/bench_chunked.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_chunked.json").read().strip().splitlines()[-1])
print("value %.0f ms %.2f e2e %.0f frac %.3f"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]), d["gpu_launches"], d.get("gpu_launches_e2e"), d["clocks"])
PY
