#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -q > gpurun_out/f3_tests.txt 2>&1; echo "pytest rc $?" >> gpurun_out/f3_tests.txt
tail -8 gpurun_out/f3_tests.txt
ICH_B200_FUSE_HEAD=1 timeout 80 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/f3_launches_fused.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/f3_ncu.log 2>&1
echo "ncu rc $?"; wc -l gpurun_out/f3_launches_fused.csv
