mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "thread_per_problem or select_generic or ladder" > gpurun_out/pytest_tpp.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_tpp.log
for d in "3 1" "4 2" "5 1"; do set -- $d
  for thr in 0 1099511627776; do
    HOP_TPP_MIN_BATCH=$thr python tools/prof_s2.py --d $1 --m $2 --N 128 --B 524288 2>&1 | tail -1
  done
done
HOP_TPP_MIN_BATCH=0 python tools/prof_s2.py --d 4 --m 2 --N 128 --B 16384 2>&1 | tail -1
HOP_TPP_MIN_BATCH=1099511627776 python tools/prof_s2.py --d 4 --m 2 --N 128 --B 16384 2>&1 | tail -1
HOP_TPP_MIN_BATCH=0 python tools/prof_s2.py --d 4 --m 2 --N 128 --B 4096 2>&1 | tail -1
HOP_TPP_MIN_BATCH=1099511627776 python tools/prof_s2.py --d 4 --m 2 --N 128 --B 4096 2>&1 | tail -1
HOP_TPP_MIN_BATCH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_select_generic_tpp -c 1 -o gpurun_out/full_tpp4 python tools/prof_s2.py --d 4 --m 2 --N 128 --B 131072 --reps 1 > gpurun_out/ncu_tpp4.log 2>&1; echo "ncu rc=$?"
HOP_TPP_MIN_BATCH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_select_generic_tpp -c 1 -o gpurun_out/full_tpp5 python tools/prof_s2.py --d 5 --m 1 --N 128 --B 131072 --reps 1 > gpurun_out/ncu_tpp5.log 2>&1; echo "ncu rc=$?"
