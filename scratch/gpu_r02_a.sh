#!/bin/bash
# round 2, GPU call A: full parity suite, smoke, bench lines of all six workloads, the on-box cuDNN incumbent, per-layer table, launch list
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r02_box_a.txt
python -m pytest tests -m gpu -q -x --deselect tests/test_parity_r2.py > $O/r02_pytest_a_old.log 2>&1; echo "rc=$?" >> $O/r02_pytest_a_old.log
python -m pytest tests/test_parity_r2.py -m gpu -q > $O/r02_pytest_a_new.log 2>&1; echo "rc=$?" >> $O/r02_pytest_a_new.log
python __graft_entry__.py smoke > $O/r02_smoke_a.log 2>&1; echo "rc=$?" >> $O/r02_smoke_a.log
for c in cfg3 cfg1 cfg2 cfg4g cfg4l cfg5; do
  python bench.py --config $c --steps 10 --warmup 3 > $O/r02a_bench_$c.json 2> $O/r02a_bench_$c.err
done
for c in cfg3 cfg1 cfg2 cfg4g cfg4l cfg5; do for m in bf16 tf32; do
  timeout 300 python bench.py --impl cudnn --cudnn-mode $m --config $c --steps 5 --warmup 3 > $O/r02a_cudnn_${c}_$m.json 2> $O/r02a_cudnn_${c}_$m.err
done; done
python scratch/bench_conv.py > $O/r02a_conv_layers.txt 2>&1
ICH_B200_FOLD_EVAL_BN=1 python bench.py --config cfg5 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02a_bench_cfg5_fold.json 2> $O/r02a_bench_cfg5_fold.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv --log-file $O/r02a_launches_cfg3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r02a_ncu.log 2>&1
ls -la $O | tail -50
