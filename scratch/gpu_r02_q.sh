#!/bin/bash
# round 2, GPU call Q: per-plane staging (kd-split) for the 64-wide cout blocks, A/B
mkdir -p gpurun_out; O=gpurun_out
for v in 0 1 0 1; do
  echo "== ICH_TC_KDS64=$v" >> $O/r02q_kds64.txt
  ICH_TC_KDS64=$v timeout 100 python scratch/bench_conv.py d1.c2,u1.c1,u1.c2 5 2>&1 | grep -v total >> $O/r02q_kds64.txt
done
cat $O/r02q_kds64.txt
