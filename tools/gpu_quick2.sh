#!/bin/bash
mkdir -p gpurun_out
for mb in 3 4 2; do
  HOP_MMA_MINBLOCKS=$mb timeout 300 python tools/prof_s1.py --B 65536 --reps 3 > gpurun_out/quick_prof_mb$mb.log 2>&1; echo "minblocks $mb: $(tail -1 gpurun_out/quick_prof_mb$mb.log)"
done
