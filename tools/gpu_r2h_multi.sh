#!/bin/bash
# Round-2 closing multi-GPU pass: BASELINE configuration 4 (16 384 quadrotor HOP-DDP solves, strong scaling) under torchrun on G GPUs of
# one box, after the one-pass backward gains.  usage: gpu_r2h_multi.sh G
G=${1:-8}
mkdir -p gpurun_out
P=$((29600 + G))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $P tests/run_configs.py --configs 4 > gpurun_out/r2h_config4_${G}gpu.jsonl 2> gpurun_out/r2h_config4_${G}gpu.err; echo "configs rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2h_config4_${G}gpu.jsonl"):
    if not l.startswith("{"): continue
    d=json.loads(l)
    print("config4", d["n_gpus"], "gpus", "%.4f s"%d["device_s"], d["phase_seconds_rank0"], {k:d["parity_vs_oracle"][k]["T_hist_identical"] for k in ("fast","exact")}, "of", d["parity_vs_oracle"]["checked"])
PY
tail -2 gpurun_out/r2h_config4_${G}gpu.err | cut -c1-300
