#!/bin/bash
# round 2, GPU call M (2 GPUs): the driver's multi-GPU command on the final build: graph replay incl. NCCL by default, clean exit
mkdir -p gpurun_out; O=gpurun_out
run() { name=$1; shift; SECONDS=0; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus 2 "$@" > $O/r02m_$name.json 2> $O/r02m_$name.err; echo "$name rc=$? wall=${SECONDS}s"; tail -c 150 $O/r02m_$name.json; echo; }
run cfg3_n2 --steps 20 --warmup 5
run cfg3_n2_ref --impl reference --steps 3 --warmup 1
run cfg4g_n2_global --config cfg4g --steps 10 --warmup 3
ls $O | grep r02m
