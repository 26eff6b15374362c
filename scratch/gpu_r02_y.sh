#!/bin/bash
# round 2, GPU call Y: verification of the tree with the shared accumulator ring (parity suite incl. both ring layouts, smoke, the driver's
# two bench commands) + the remaining ablations of the streaming kernel (MMA phase alone in situ)
mkdir -p gpurun_out; O=gpurun_out
P=label-efficient-volumetric-deep-semantic-segmentation-of-ich_b200/ich_b200
timeout 700 python -m pytest tests -m gpu -q > $O/r02y_pytest.log 2>&1; echo "rc=$?" >> $O/r02y_pytest.log
timeout 120 python __graft_entry__.py smoke > $O/r02y_smoke.log 2>&1; echo "rc=$?" >> $O/r02y_smoke.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02y_bench_cfg3.json 2> $O/r02y_bench_cfg3.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02y_reference_cfg3.json 2> $O/r02y_reference_cfg3.err
{
echo "== release build"; timeout 100 python scratch/bench_conv.py 2>&1
for dbg in 2 6; do
  echo "== ablation build, ICH_TC_DBG=$dbg (2 = no TMA loads, 6 = no TMA loads and no epilogue math / stores)"
  ICH_B200_LIB=$P/libich_b200_dbg.so ICH_TC_DBG=$dbg timeout 100 python scratch/bench_conv.py d0.c2,u2.c1,u2.c2 5 2>&1 | grep -v total | sed 's/| wgrad.*//'
done
} > $O/r02y_conv_layers.txt 2>&1
tail -3 $O/r02y_pytest.log; tail -2 $O/r02y_smoke.log; cat $O/r02y_conv_layers.txt
