#!/bin/bash
# round 2, GPU call R: streaming kernel with 2-CTA clusters + TMA-multicast weights on the mid-resolution layers
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02r_cluster.txt
for cfg in "ICH_TC_STREAM=2 ICH_TC_STREAM_CLUSTER=2" "ICH_TC_STREAM=2 ICH_TC_STREAM_CLUSTER=4"; do
  echo "=== $cfg" >> $O/r02r_cluster.txt
  env $cfg timeout 120 python scratch/check_stream_cluster.py >> $O/r02r_cluster.txt 2>&1; echo "rc=$?" >> $O/r02r_cluster.txt
done
cat $O/r02r_cluster.txt
