#!/bin/bash
# round 2, last call: the driver's 4-GPU bench command on the final tree (graph replay incl. the captured NCCL all-reduce)
mkdir -p gpurun_out; O=gpurun_out
S=$SECONDS
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --steps 20 --warmup 5 > $O/r02_final_bench_n4.json 2> $O/r02_final_bench_n4.err
echo "rc=$? wall=$((SECONDS-S))s" >> $O/r02_final_bench_n4.err
tail -1 $O/r02_final_bench_n4.json | cut -c1-400; tail -2 $O/r02_final_bench_n4.err
