#!/bin/bash
# round 2, GPU call J: final build -- parity suite, smoke, the driver's own commands, every config
mkdir -p gpurun_out; O=gpurun_out
timeout 700 python -m pytest tests -m gpu -q > $O/r02_pytest_j.log 2>&1; echo "rc=$?" >> $O/r02_pytest_j.log
timeout 120 python __graft_entry__.py smoke > $O/r02_smoke_j.log 2>&1; echo "rc=$?" >> $O/r02_smoke_j.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02j_bench_cfg3.json 2> $O/r02j_bench_cfg3.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02j_reference_cfg3.json 2> $O/r02j_reference_cfg3.err
for c in cfg1 cfg2 cfg4g cfg4l cfg5; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 > $O/r02j_bench_$c.json 2> $O/r02j_bench_$c.err
done
timeout 200 python bench.py --config cfg5 --steps 10 --warmup 3 --graph 0 --no-cpu-baseline > $O/r02j_bench_cfg5_eager.json 2>> $O/r02j_bench.err
timeout 200 python scratch/bench_conv.py > $O/r02j_conv_layers.txt 2>&1
ls $O | grep r02j
