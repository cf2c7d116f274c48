mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for c in Segway_Balance Cartpole_SwingUp; do for B in 4096 131072; do
  echo -n "lanes "; HOP_FUSED_LANES=1 python tools/prof_small.py --case $c --B $B 2>&1 | tail -1
done; done
HOP_TPP_MIN_BATCH=1099511627776 python tools/prof_s2.py --d 5 --m 1 --N 128 --B 262144 2>&1 | tail -1
HOP_TPP_MIN_BATCH=1099511627776 python tools/prof_s2.py --d 4 --m 2 --N 128 --B 262144 2>&1 | tail -1
HOP_TPP_MIN_BATCH=1099511627776 python tools/prof_s2.py --d 5 --m 1 --N 128 --B 2048 --reps 10 2>&1 | tail -1
timeout 900 python tests/run_configs.py --configs 3 > gpurun_out/cfg3_prefetch.jsonl 2> gpurun_out/cfg3.err; python -c "
import json
for l in open('gpurun_out/cfg3_prefetch.jsonl'):
    d=json.loads(l); print(d['config'], d['device_s'], d['phase_seconds_rank0'])"
