for mb in 2 3 4; do for v in 0 1; do
  echo -n "MINB=$mb U2=$v "; HOP_MMA_MINBLOCKS=$mb HOP_PIPE_U2=$v python tools/prof_s1.py --B 65536 --reps 3 2>&1 | tail -1
done; done
