#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; rm -f $O/r02t_stream_nb32.txt
for cfg in "ICH_TC_STREAM=2 ICH_TC_STREAM_NB=32" "ICH_TC_STREAM=2 ICH_TC_STREAM_NB=32 ICH_TC_STREAM_ISSUERS=3"; do
  echo "=== $cfg" >> $O/r02t_stream_nb32.txt
  env $cfg timeout 120 python scratch/check_stream_cluster.py >> $O/r02t_stream_nb32.txt 2>&1; echo "rc=$?" >> $O/r02t_stream_nb32.txt
done
cat $O/r02t_stream_nb32.txt
