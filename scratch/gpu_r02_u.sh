#!/bin/bash
# round 2, GPU call U: ncu launch list of the default bench on the final tree (eager launches so that every kernel is a separate launch)
mkdir -p gpurun_out; O=gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv --log-file $O/r02u_launches_cfg3.csv python bench.py --steps 1 --warmup 3 --graph 0 --no-cpu-baseline > $O/r02u_ncu.log 2>&1
ls -la $O | grep r02u
