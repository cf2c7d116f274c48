#!/bin/bash
# Round-2 GPU pass A (run under gpurun): GPU tests, smoke, bench (fast / exact / reference arm).  Output: gpurun_out/r2a_*
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc >> gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/r2a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/r2a_smoke.log
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
timeout 600 python bench.py --mode exact --steps 2 > gpurun_out/r2a_bench_exact.json 2> gpurun_out/r2a_bench_exact.err; echo "bench exact rc=$?"; cut -c1-400 gpurun_out/r2a_bench_exact.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; echo "bench ref rc=$?"; cut -c1-300 gpurun_out/r2a_bench_ref.json
