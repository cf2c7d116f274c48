#!/bin/bash
# Round-2 final 1-GPU pass (r2h: after the warp-specialised selection pipeline and the one-pass backward gains): GPU tests, smoke, bench (fast, exact, reference arm), the five BASELINE configurations,
# ncu launch list of the bench command and one --set full capture of the headline kernel.  Output: gpurun_out/r2h_*
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2h_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2h_smoke.log
timeout 600 python bench.py > gpurun_out/r2h_bench_1gpu.json 2> gpurun_out/r2h_bench_1gpu.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/r2h_bench_1gpu.json
timeout 600 python bench.py --impl reference > gpurun_out/r2h_bench_reference_arm.json 2> gpurun_out/r2h_bench_reference_arm.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r2h_bench_reference_arm.json
timeout 600 python bench.py --mode exact --steps 2 > gpurun_out/r2h_bench_exact_mode.json 2> /dev/null; echo "exact rc=$?"; cut -c1-200 gpurun_out/r2h_bench_exact_mode.json
timeout 1800 python tests/run_configs.py > gpurun_out/r2h_configs_1gpu.jsonl 2> gpurun_out/r2h_configs.err; echo "configs rc=$?"; tail -2 gpurun_out/r2h_configs.err | cut -c1-200
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-legs > gpurun_out/r2h_bench_nocpu.json 2> gpurun_out/r2h_bench_nocpu.err && \
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-legs > gpurun_out/r2h_ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 300 python tools/prof_s1.py --B 65536 --reps 2 > gpurun_out/r2h_prof_plain.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_select_fused_mma -s 1 -c 1 -o gpurun_out/r2h_prof_select python tools/prof_s1.py --B 65536 --reps 2 > gpurun_out/r2h_ncu_full.log 2>&1
echo "ncu full rc=$?"
