#!/bin/bash
# Round-2 multi-GPU pass: bench.py and BASELINE configurations 4, 5 under torchrun on G GPUs of one box.  usage: gpu_r2_multi.sh G
G=${1:-8}
mkdir -p gpurun_out
P=$((29500 + G))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $P bench.py --gpus $G --steps 5 --warmup 3 > gpurun_out/r2_bench_${G}gpu.json 2> gpurun_out/r2_bench_${G}gpu.err; echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_${G}gpu.json | cut -c1-330; tail -2 gpurun_out/r2_bench_${G}gpu.err | cut -c1-300
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((P+20)) tests/run_configs.py --configs 4,5 > gpurun_out/r2_configs_${G}gpu.jsonl 2> gpurun_out/r2_configs_${G}gpu.err; echo "configs rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2_configs_${G}gpu.jsonl"):
    if not l.startswith("{"): continue
    d=json.loads(l)
    if d["config"]==4: print("config4", d["n_gpus"], "gpus", "%.4f s"%d["device_s"], d["phase_seconds_rank0"], {k:d["parity_vs_oracle"][k]["T_hist_identical"] for k in ("fast","exact")}, "of", d["parity_vs_oracle"]["checked"])
    else: print(d["what"], d["n_gpus"], "gpus %.3g solves/s"%d["solves_per_s"], "tiles", d["tiles_per_rank"])
PY
tail -2 gpurun_out/r2_configs_${G}gpu.err | cut -c1-300
