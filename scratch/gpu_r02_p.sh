#!/bin/bash
# round 2, GPU call P: ablations of the slab kernel on the mid-resolution layers (what keeps the MMA unit only 41-62 % busy?)
mkdir -p gpurun_out; O=gpurun_out
export ICH_B200_LIB=$PWD/label-efficient-volumetric-deep-semantic-segmentation-of-ich_b200/ich_b200/libich_b200_dbg.so
for d in 0 1 2 4 3 6; do
  echo "== ICH_TC_DBG=$d" >> $O/r02p_ablate_slab.txt
  ICH_TC_DBG=$d timeout 100 python scratch/bench_conv.py d1.c1,u1.c2,u1.c1,d2.c2 3 2>&1 | grep -v total >> $O/r02p_ablate_slab.txt
done
cat $O/r02p_ablate_slab.txt
