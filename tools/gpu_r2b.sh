#!/bin/bash
# Round-2 GPU pass B: GPU tests, e2e A/B of the host-buffer entry, BASELINE configurations 1-5.  Output: gpurun_out/r2b_*
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2b_pytest.log | cut -c1-300
for v in "prio8:" "serial8:HOP_HOST_SERIAL_PREP=1" "prio4:HOP_HOST_CHUNKS=4" "serial4:HOP_HOST_SERIAL_PREP=1,HOP_HOST_CHUNKS=4" "prio16:HOP_HOST_CHUNKS=16"; do
  name=${v%%:*}; envs=$(echo ${v#*:} | tr ',' ' ')
  env $envs timeout 300 python tools/e2e_ab.py > gpurun_out/r2b_e2e_$name.log 2>&1; echo "$name: $(tail -1 gpurun_out/r2b_e2e_$name.log)"
done
timeout 600 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; cut -c1-500 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
timeout 1500 python tests/run_configs.py --configs 1,2,3,4 > gpurun_out/r2b_configs_1gpu.jsonl 2> gpurun_out/r2b_configs.err; echo "configs rc=$?"; cut -c1-1500 gpurun_out/r2b_configs_1gpu.jsonl; tail -3 gpurun_out/r2b_configs.err
timeout 1500 python tests/run_configs.py --configs 5 > gpurun_out/r2b_config5_1gpu.jsonl 2> gpurun_out/r2b_config5.err; echo "config5 rc=$?"; cut -c1-700 gpurun_out/r2b_config5_1gpu.jsonl; tail -3 gpurun_out/r2b_config5.err
