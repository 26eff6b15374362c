#!/bin/bash
# round 2, GPU call B: parity suite with the new rows (staging, window driver, segement_volume, MLPHead, GatedUNet, graph), A/B of the
# fused-head backward kernels, CUDA-graph replay
mkdir -p gpurun_out; O=gpurun_out
python -m pytest tests -m gpu -q > $O/r02_pytest_b.log 2>&1; echo "rc=$?" >> $O/r02_pytest_b.log
python __graft_entry__.py smoke > $O/r02_smoke_b.log 2>&1; echo "rc=$?" >> $O/r02_smoke_b.log
python bench.py --steps 20 --warmup 5 > $O/r02b_bench_cfg3.json 2> $O/r02b_bench_cfg3.err
ICH_HEAD_APPLY_U=8 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02b_bench_cfg3_applyU8.json 2> $O/r02b_bench_cfg3_applyU8.err
python bench.py --steps 20 --warmup 5 --graph 1 --no-cpu-baseline > $O/r02b_bench_cfg3_graph.json 2> $O/r02b_bench_cfg3_graph.err
for c in cfg1 cfg2 cfg4g; do
  python bench.py --config $c --steps 10 --warmup 3 --graph 1 --no-cpu-baseline > $O/r02b_bench_${c}_graph.json 2> $O/r02b_bench_${c}_graph.err
done
python bench.py --config cfg5 --steps 10 --warmup 3 > $O/r02b_bench_cfg5.json 2> $O/r02b_bench_cfg5.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02b_reference_cfg3.json 2> $O/r02b_reference_cfg3.err
ICH_TC_STREAM=2 python scratch/bench_conv.py d1.c1,d1.c2,u1.c1,u1.c2,d2.c2,u0.c2 > $O/r02b_conv_layers_stream2.txt 2>&1
ls -la $O | tail -20
