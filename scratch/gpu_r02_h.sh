#!/bin/bash
# round 2, GPU call H: transposed-conv backward without the re-pack (A/B), fused-head apply shape (A/B), parity suite
mkdir -p gpurun_out; O=gpurun_out
timeout 700 python -m pytest tests -m gpu -q -x > $O/r02_pytest_h.log 2>&1; echo "rc=$?" >> $O/r02_pytest_h.log
for v in 1 0 1 0; do
  ICH_B200_CONVT_DIRECT=$v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02h_bench_cfg3_direct${v}_$RANDOM.json 2>> $O/r02h_bench.err
done
for u in 2 4; do
  ICH_HEAD_BWD_U=$u timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02h_bench_cfg3_applyU${u}.json 2>> $O/r02h_bench.err
done
for v in 1 0; do
  ICH_B200_CONVT_DIRECT=$v timeout 200 python bench.py --config cfg2 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02h_bench_cfg2_direct$v.json 2>> $O/r02h_bench.err
done
ls $O | grep r02h
