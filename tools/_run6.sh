mkdir -p gpurun_out
for v in 0 1; do
  HOP_PIPE_U2=$v python tools/prof_s1.py --B 65536 --reps 3 2>&1 | tail -1
done
HOP_PIPE_U2=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "s1_from or fast_and_exact or full_size or golden" > gpurun_out/pytest_u2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_u2.log
HOP_PIPE_U2=1 timeout 600 python bench.py --steps 3 > gpurun_out/bench_u2.json 2> gpurun_out/bench_u2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_u2.json").read().strip().splitlines()[-1])
print("value %.0f ms %.2f e2e %.0f frac %.3f"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]), d["cpu_baseline"]["parity_on_sample"])
PY
