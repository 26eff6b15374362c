#!/bin/bash
# round 2, GPU call D (2 GPUs): data-parallel step with and without CUDA-graph replay, cross-rank contrastive set, local contrastive,
# inference sharded by volume and by window
mkdir -p gpurun_out; O=gpurun_out
run() { name=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus 2 "$@" > $O/r02d_$name.json 2> $O/r02d_$name.err; tail -c 300 $O/r02d_$name.json; echo; }
python -m pytest tests/test_parity_r2.py -m gpu -q -k "graphed or segement or window_driver" > $O/r02_pytest_d.log 2>&1; echo "rc=$?" >> $O/r02_pytest_d.log
run cfg3_n2 --steps 20 --warmup 5
run cfg3_n2_graph --steps 20 --warmup 5 --graph 1
ICH_B200_GLOBAL_NCE=1 run cfg4g_n2_global --config cfg4g --steps 10 --warmup 3
run cfg4l_n2 --config cfg4l --steps 10 --warmup 3
run cfg5_n2_volume --config cfg5 --steps 10 --warmup 3
run cfg5_n2_window --config cfg5 --steps 10 --warmup 3 --shard window
timeout 600 ncu --set full --clock-control none -k regex:'bn_head|space_to_depth' --launch-skip 8 -c 5 -o $O/r02d_head_s2d python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r02d_ncu.log 2>&1
ncu -i $O/r02d_head_s2d.ncu-rep --page raw --csv > $O/r02d_head_s2d_raw.csv 2>/dev/null
ls $O | grep r02d
