mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_select_fused_mma -c 1 -o gpurun_out/full_s1_final python tools/prof_s1.py --B 65536 --reps 1 > gpurun_out/ncu_s1.log 2>&1; echo "ncu rc=$?"
timeout 1500 python tests/run_configs.py > gpurun_out/configs_1gpu.jsonl 2> gpurun_out/configs_1gpu.err; echo "configs rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_1gpu.json").read().strip().splitlines()[-1])
print("value %.0f ms %.2f e2e %.0f frac %.3f"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]), d["clocks"], d["cpu_baseline"]["value"], d["cpu_baseline"]["parity_on_sample"]["T_star_mismatches"])
for l in open("gpurun_out/configs_1gpu.jsonl"):
    d=json.loads(l); print(d["config"], d.get("what","")[:40], d.get("device_s"), d.get("solves_per_s"), d.get("hbm_frac_per_gpu"), d.get("algorithmic_TFLOPs"))
PY
