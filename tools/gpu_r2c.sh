#!/bin/bash
# Round-2 GPU pass C: GPU tests, bench, configs 2-4 (census), then the ncu launch list of the bench command and one
# --set full capture of the headline kernel.  Output: gpurun_out/r2c_*
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2c_pytest.log | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2c_smoke.log
timeout 600 python bench.py --mode exact --steps 2 > gpurun_out/r2c_bench_exact.json 2> gpurun_out/r2c_bench_exact.err; echo "bench exact rc=$?"; cut -c1-260 gpurun_out/r2c_bench_exact.json
timeout 1500 python tests/run_configs.py --configs 2,3,4 > gpurun_out/r2c_configs_1gpu.jsonl 2> gpurun_out/r2c_configs.err; echo "configs rc=$?"; cut -c1-2500 gpurun_out/r2c_configs_1gpu.jsonl; tail -3 gpurun_out/r2c_configs.err
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err && \
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2c_ncu_list.log 2>&1
echo "bench+ncu list rc=$?"; cut -c1-300 gpurun_out/r2c_bench.json
timeout 300 python tools/prof_s1.py --B 65536 --reps 2 > gpurun_out/r2c_prof_plain.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_select_fused_mma -s 1 -c 1 -o gpurun_out/r2c_prof_select python tools/prof_s1.py --B 65536 --reps 2 > gpurun_out/r2c_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/r2c_ncu_full.log
