#!/bin/bash
# round 2, GPU call I: bias gradient of the transposed conv from the data-gradient epilogue (A/B), parity suite
mkdir -p gpurun_out; O=gpurun_out
timeout 700 python -m pytest tests -m gpu -q -x > $O/r02_pytest_i.log 2>&1; echo "rc=$?" >> $O/r02_pytest_i.log
for v in 1 0 1 0; do
  ICH_B200_DGRAD_COLSUM=$v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02i_bench_cfg3_colsum${v}_$RANDOM.json 2>> $O/r02i_bench.err
done
for v in 1 0; do
  ICH_B200_DGRAD_COLSUM=$v timeout 200 python bench.py --config cfg2 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02i_bench_cfg2_colsum$v.json 2>> $O/r02i_bench.err
done
ls $O | grep r02i
