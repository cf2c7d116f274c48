mkdir -p gpurun_out
python tools/prof_s1.py --B 16384 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_select_fused_mma -c 1 -o gpurun_out/full_s1 python tools/prof_s1.py --B 16384 --reps 1 > gpurun_out/ncu_s1.log 2>&1; echo "ncu rc=$?"
