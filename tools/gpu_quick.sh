#!/bin/bash
# quick A/B of the headline kernel: pipeline timing (tools/prof_s1.py) + the S1 / mode tests
mkdir -p gpurun_out
timeout 300 python tools/prof_s1.py --B 65536 --reps 3 > gpurun_out/quick_prof.log 2>&1; tail -1 gpurun_out/quick_prof.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-legs > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/quick_bench.json").read().strip().splitlines()[-1]); print("value %.0f ms %.3f e2e %.0f frac %.4f"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
PY
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "s1_from_x0 or fast_and_gj or exact_mode_is_bit or nonzero_affine or select_fused_matches or full_size" > gpurun_out/quick_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/quick_pytest.log | cut -c1-300
