#!/bin/bash
# round 2, GPU call C: parity suite, A/B of the fused-head backward shapes, ncu --set full of the HBM-bound outliers
mkdir -p gpurun_out; O=gpurun_out
python -m pytest tests -m gpu -q > $O/r02_pytest_c.log 2>&1; echo "rc=$?" >> $O/r02_pytest_c.log
for u in 4 8 4 8; do
  ICH_HEAD_BWD_U=$u python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02c_bench_cfg3_U${u}_$RANDOM.json 2>> $O/r02c_bench.err
done
python bench.py --config cfg4l --steps 10 --warmup 3 --no-cpu-baseline > $O/r02c_bench_cfg4l.json 2>> $O/r02c_bench.err
timeout 900 ncu --set full --clock-control none -k regex:'bn_head|space_to_depth|cin1|seg_loss|permute5' -c 10 -o $O/r02c_hbm_outliers python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r02c_ncu.log 2>&1
ncu -i $O/r02c_hbm_outliers.ncu-rep --page raw --csv > $O/r02c_hbm_outliers_raw.csv 2>/dev/null
ls -la $O | tail -12
