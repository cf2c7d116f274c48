mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench8 rc=$?"; tail -c 600 gpurun_out/bench_8gpu.json | head -c 400; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tests/run_configs.py --configs 4,5 > gpurun_out/configs_8gpu.jsonl 2> gpurun_out/configs_8gpu.err; echo "cfg8 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_8gpu.json","gpurun_out/bench_2gpu.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, "value %.0f ms %.2f e2e %.0f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), d["n_gpus"])
for l in open("gpurun_out/configs_8gpu.jsonl"):
    d=json.loads(l); print(d["config"], d.get("what","")[:40], d.get("device_s"), d.get("solves_per_s"), d.get("n_gpus"))
PY
