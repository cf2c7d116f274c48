#!/bin/bash
# Round-2 GPU pass D: GPU tests, diagonal fast path A/B on S2 (d = 12, 13), config 5, ncu launch list of the bench command.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2d_pytest.log | cut -c1-300
for d in 13 12; do
  for v in "diag:" "sweep:HOP_GENERIC_NODIAG=1"; do
    name=${v%%:*}; envs=${v#*:}
    env $envs timeout 300 python tools/prof_s2.py --d $d --m 4 --N 128 --B 65536 --reps 3 > gpurun_out/r2d_s2_d${d}_$name.log 2>&1; echo "d=$d $name: $(tail -1 gpurun_out/r2d_s2_d${d}_$name.log)"
  done
done
timeout 1500 python tests/run_configs.py --configs 5 > gpurun_out/r2d_config5_1gpu.jsonl 2> gpurun_out/r2d_config5.err; echo "config5 rc=$?"; python - <<'PY'
import json
for l in open("gpurun_out/r2d_config5_1gpu.jsonl"):
    d=json.loads(l); print(d["what"], "%.3g solves/s"%d["solves_per_s"], "TF %.2f"%d["algorithmic_TFLOPs"], "hbm %.3f"%d["hbm_frac_per_gpu"], d["parity_vs_oracle"]["max_rel_J"], d["parity_vs_oracle"]["T_star_identical"], "x%.0f vs cpu bf"%d["speedup_vs_cpu_bruteforce_port_all_cores"])
PY
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-legs > gpurun_out/r2d_bench_nocpu.json 2> gpurun_out/r2d_bench_nocpu.err && \
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-legs > gpurun_out/r2d_ncu_list.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/r2d_launches.csv
