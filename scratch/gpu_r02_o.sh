#!/bin/bash
# round 2, GPU call O: a third MMA issuer warp (slab kernel by default when an item has >= 3 tiles; streaming kernel opt-in), A/B + parity suite
mkdir -p gpurun_out; O=gpurun_out
timeout 700 python -m pytest tests -m gpu -q -x > $O/r02_pytest_o.log 2>&1; echo "rc=$?" >> $O/r02_pytest_o.log
ICH_TC_ISSUERS=2 timeout 200 python scratch/bench_conv.py > $O/r02o_conv_layers_i2.txt 2>&1
ICH_TC_ISSUERS=3 timeout 200 python scratch/bench_conv.py > $O/r02o_conv_layers_i3.txt 2>&1
ICH_TC_ISSUERS=3 ICH_TC_STREAM_ISSUERS=3 timeout 200 python scratch/bench_conv.py > $O/r02o_conv_layers_i3s3.txt 2>&1
for v in 2 3 2 3; do
  ICH_TC_ISSUERS=$v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02o_bench_cfg3_i${v}_$RANDOM.json 2>> $O/r02o_bench.err
done
ICH_TC_ISSUERS=3 ICH_TC_STREAM_ISSUERS=3 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02o_bench_cfg3_i3s3.json 2>> $O/r02o_bench.err
for v in 2 3; do
  ICH_TC_ISSUERS=$v timeout 200 python bench.py --config cfg2 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02o_bench_cfg2_i$v.json 2>> $O/r02o_bench.err
done
ls $O | grep r02o
