mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
python tools/prof_s2.py --d 4 --m 2 --N 128 --B 262144 > gpurun_out/s2_d4.json 2>&1
python tools/prof_s2.py --d 5 --m 1 --N 128 --B 262144 > gpurun_out/s2_d5.json 2>&1
python tools/prof_s2.py --d 13 --m 4 --N 128 --B 32768 > gpurun_out/s2_d13.json 2>&1
cat gpurun_out/s2_d*.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_select_generic -c 1 -o gpurun_out/full_d4 python tools/prof_s2.py --d 4 --m 2 --N 128 --B 131072 --reps 1 > gpurun_out/ncu_d4.log 2>&1; echo "ncu d4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_select_generic -c 1 -o gpurun_out/full_d13 python tools/prof_s2.py --d 13 --m 4 --N 128 --B 16384 --reps 1 > gpurun_out/ncu_d13.log 2>&1; echo "ncu d13 rc=$?"
ls -la gpurun_out
