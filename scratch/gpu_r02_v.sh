#!/bin/bash
# round 2, GPU call V: last verification of the final tree (parity suite, smoke, the driver's two bench commands) + ncu --set full of the
# full-resolution conv layers (streaming kernel fwd / dgrad, weight gradient) for the tensor-pipe utilisation figures of this round
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q > $O/r02v_pytest.log 2>&1; echo "rc=$?" >> $O/r02v_pytest.log
timeout 120 python __graft_entry__.py smoke > $O/r02v_smoke.log 2>&1; echo "rc=$?" >> $O/r02v_smoke.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02v_bench_cfg3.json 2> $O/r02v_bench_cfg3.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02v_reference_cfg3.json 2> $O/r02v_reference_cfg3.err
timeout 400 ncu --set full --clock-control none -k regex:'conv_tc' --launch-skip 0 -c 27 -o $O/r02v_fullres python scratch/bench_conv.py u2.c1,u2.c2,d0.c2 1 > $O/r02v_ncu.log 2>&1
ncu -i $O/r02v_fullres.ncu-rep --page raw --csv > $O/r02v_fullres_raw.csv 2>/dev/null
rm -f $O/r02v_fullres.ncu-rep
ls -la $O | grep r02v
