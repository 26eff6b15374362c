#!/bin/bash
# round 2, GPU call K (8 GPUs): the driver's scaling command on the final build (graph replay incl. NCCL by default), exit behaviour
mkdir -p gpurun_out; O=gpurun_out
run() { name=$1; shift; SECONDS=0; timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus 8 "$@" > $O/r02k_$name.json 2> $O/r02k_$name.err; echo "$name rc=$? wall=${SECONDS}s"; tail -c 150 $O/r02k_$name.json; echo; }
run cfg3_n8 --steps 20 --warmup 5
ls $O | grep r02k
