#!/bin/bash
# round 2, GPU call S: final tree -- parity suite, smoke, the driver's bench commands
mkdir -p gpurun_out; O=gpurun_out
timeout 800 python -m pytest tests -m gpu -q > $O/r02_pytest_s.log 2>&1; echo "rc=$?" >> $O/r02_pytest_s.log
timeout 120 python __graft_entry__.py smoke > $O/r02_smoke_s.log 2>&1; echo "rc=$?" >> $O/r02_smoke_s.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02s_bench_cfg3.json 2> $O/r02s_bench_cfg3.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02s_reference_cfg3.json 2> $O/r02s_reference_cfg3.err
ls $O | grep "r02s\|_s.log"
