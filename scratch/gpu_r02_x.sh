#!/bin/bash
# round 2, GPU call X: ONE accumulator ring shared by the tiles of an item in the plane-streaming kernel (ICH_TC_STREAM_CRING, default 1):
# the kd-fold MMA splits in two for 2 of every 16 planes instead of 2 of every 4.  A/B per layer + parity suite + whole step.
mkdir -p gpurun_out; O=gpurun_out/r02x_stream_cring.txt
{
echo "== CRING=0 (one 4-slot ring per tile)"; ICH_TC_STREAM_CRING=0 timeout 100 python scratch/bench_conv.py d0.c2,u2.c1,u2.c2,d1.c2 5 2>&1 | sed 's/| wgrad.*//'
echo "== CRING=1 (one ring of slots x T cells)"; timeout 100 python scratch/bench_conv.py d0.c2,u2.c1,u2.c2,d1.c2 5 2>&1 | sed 's/| wgrad.*//'
} > $O 2>&1
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/r02x_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02x_pytest.log
ICH_TC_STREAM_CRING=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02x_bench_cring0.json 2> gpurun_out/r02x_bench_cring0.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02x_bench_cring1.json 2> gpurun_out/r02x_bench_cring1.err
cat $O; tail -4 gpurun_out/r02x_pytest.log
