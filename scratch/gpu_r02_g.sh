#!/bin/bash
# round 2, GPU call G (8 GPUs): every multi-GPU row of SURVEY section 8e on hardware.  Every command is bounded by `timeout`.
mkdir -p gpurun_out; O=gpurun_out
run() { name=$1; shift; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus 8 "$@" > $O/r02g_$name.json 2> $O/r02g_$name.err; echo "$name rc=$?"; tail -c 200 $O/r02g_$name.json; echo; }
run cfg3_n8 --steps 20 --warmup 5 --graph 0
run cfg3_n8_graph --steps 20 --warmup 5 --graph 1
ICH_B200_GLOBAL_NCE=1 run cfg4g_n8_global --config cfg4g --steps 10 --warmup 3 --graph 0
run cfg4l_n8 --config cfg4l --steps 10 --warmup 3
run cfg5_n8_volume --config cfg5 --steps 10 --warmup 3
run cfg5_n8_window --config cfg5 --steps 10 --warmup 3 --shard window
ls $O | grep r02g
