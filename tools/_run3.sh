mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for B in 32 128 512 1024 2048 8192; do
  for thr in 0 1099511627776; do
    echo -n "thr=$thr "; HOP_TPP_MIN_BATCH=$thr python tools/prof_s2.py --d 4 --m 2 --N 128 --B $B --reps 20 2>&1 | tail -1
  done
done
for B in 32 512 2048; do
  for thr in 0 1099511627776; do
    echo -n "thr=$thr "; HOP_TPP_MIN_BATCH=$thr python tools/prof_s2.py --d 5 --m 1 --N 128 --B $B --reps 20 2>&1 | tail -1
    echo -n "thr=$thr "; HOP_TPP_MIN_BATCH=$thr python tools/prof_s2.py --d 3 --m 1 --N 128 --B $B --reps 20 2>&1 | tail -1
  done
done
