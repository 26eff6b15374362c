#!/bin/bash
# round 2, GPU call E: kd-split / 128-wide cout blocks in the slab kernel (A/B against ICH_TC_NB128=0), parity suite, ncu of the fused-head / re-pack kernels
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q > $O/r02_pytest_e.log 2>&1; echo "rc=$?" >> $O/r02_pytest_e.log
timeout 200 python scratch/bench_conv.py > $O/r02e_conv_layers_nb128.txt 2>&1
ICH_TC_NB128=0 timeout 200 python scratch/bench_conv.py d2.c2,u0.c1,u0.c2,bt.c2,u1.c1 > $O/r02e_conv_layers_nb64.txt 2>&1
for v in 1 0 1 0; do
  ICH_TC_NB128=$v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02e_bench_cfg3_nb128_${v}_$RANDOM.json 2>> $O/r02e_bench.err
done
for v in 1 0; do
  ICH_TC_NB128=$v timeout 200 python bench.py --config cfg2 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02e_bench_cfg2_nb128_$v.json 2>> $O/r02e_bench.err
  ICH_TC_NB128=$v timeout 200 python bench.py --config cfg4l --steps 10 --warmup 3 --no-cpu-baseline > $O/r02e_bench_cfg4l_nb128_$v.json 2>> $O/r02e_bench.err
done
timeout 400 ncu --set full --clock-control none -k regex:'bn_head|space_to_depth' --launch-skip 8 -c 5 -o $O/r02e_head_s2d python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r02e_ncu.log 2>&1
ncu -i $O/r02e_head_s2d.ncu-rep --page raw --csv > $O/r02e_head_s2d_raw.csv 2>/dev/null
ls $O | grep r02e
