#!/bin/bash
# round 2, GPU call W: the Cin = 16 full-resolution layer (d0.c2: 950 / 713 TFLOP/s) -- is it bound by the drain of its output planes?
# A/B of 16 epilogue warps (ICH_TC_STREAM_EPI=16) in the plane-streaming kernel + skeleton ablation + parity suite with the variant
mkdir -p gpurun_out; O=gpurun_out/r02w_stream_epi.txt
P=label-efficient-volumetric-deep-semantic-segmentation-of-ich_b200/ich_b200
{
echo "== EPI=8 (default)";  timeout 100 python scratch/bench_conv.py d0.c2,u2.c1,u2.c2 5 2>&1 | sed 's/| wgrad.*//'
echo "== EPI=16"; ICH_TC_STREAM_EPI=16 timeout 100 python scratch/bench_conv.py d0.c2,u2.c1,u2.c2 5 2>&1 | sed 's/| wgrad.*//'
for dbg in 1 7; do
  echo "== ablation build, EPI=8, ICH_TC_DBG=$dbg"; ICH_B200_LIB=$P/libich_b200_dbg.so ICH_TC_DBG=$dbg timeout 100 python scratch/bench_conv.py d0.c2,u2.c2 5 2>&1 | grep -v total | sed 's/| wgrad.*//'
done
echo "== ablation build, EPI=16, ICH_TC_DBG=7"; ICH_TC_STREAM_EPI=16 ICH_B200_LIB=$P/libich_b200_dbg.so ICH_TC_DBG=7 timeout 100 python scratch/bench_conv.py d0.c2,u2.c2 5 2>&1 | grep -v total | sed 's/| wgrad.*//'
} > $O 2>&1
ICH_TC_STREAM_EPI=16 timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/r02w_pytest_epi16.log 2>&1; echo "rc=$?" >> gpurun_out/r02w_pytest_epi16.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02w_bench_epi8.json 2> gpurun_out/r02w_bench_epi8.err
ICH_TC_STREAM_EPI=16 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02w_bench_epi16.json 2> gpurun_out/r02w_bench_epi16.err
cat $O; tail -3 gpurun_out/r02w_pytest_epi16.log
