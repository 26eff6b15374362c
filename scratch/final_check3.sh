#!/bin/bash
mkdir -p gpurun_out
timeout 40 python -m pytest tests -m gpu -q -k "head or end_to_end" > gpurun_out/f4_tests.txt 2>&1; echo "pytest rc $?" >> gpurun_out/f4_tests.txt
tail -4 gpurun_out/f4_tests.txt
ICH_B200_FUSE_HEAD=1 timeout 40 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f4_bench_fused.json 2> gpurun_out/f4_bench_fused.err
python -c "
import json; d=json.load(open('gpurun_out/f4_bench_fused.json')); print('fused v2', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'])"
