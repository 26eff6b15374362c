#!/bin/bash
# round 2, GPU call L: new parity cases (floor pooling, wide head), cfg-4 global after the unused-skip-gradient fix
mkdir -p gpurun_out; O=gpurun_out
timeout 700 python -m pytest tests -m gpu -q > $O/r02_pytest_l.log 2>&1; echo "rc=$?" >> $O/r02_pytest_l.log
timeout 200 python bench.py --config cfg4g --steps 10 --warmup 3 --no-cpu-baseline > $O/r02l_bench_cfg4g.json 2> $O/r02l_bench_cfg4g.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02l_bench_cfg3.json 2> $O/r02l_bench_cfg3.err
ls $O | grep r02l
