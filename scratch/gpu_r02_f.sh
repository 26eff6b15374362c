#!/bin/bash
# round 2, GPU call F: parity suite (padded mid channels, kd-split, graph), cfg-1 / cfg-3 / cfg-2 lines, graph A/B, launch list
mkdir -p gpurun_out; O=gpurun_out
timeout 700 python -m pytest tests -m gpu -q > $O/r02_pytest_f.log 2>&1; echo "rc=$?" >> $O/r02_pytest_f.log
timeout 120 python __graft_entry__.py smoke > $O/r02_smoke_f.log 2>&1; echo "rc=$?" >> $O/r02_smoke_f.log
for g in 0 1 0 1; do
  timeout 200 python bench.py --steps 20 --warmup 5 --graph $g --no-cpu-baseline > $O/r02f_bench_cfg3_graph${g}_$RANDOM.json 2>> $O/r02f_bench.err
done
timeout 200 python bench.py --config cfg1 --steps 20 --warmup 5 --graph 1 > $O/r02f_bench_cfg1.json 2>> $O/r02f_bench.err
timeout 200 python bench.py --config cfg2 --steps 10 --warmup 3 --graph 1 --no-cpu-baseline > $O/r02f_bench_cfg2.json 2>> $O/r02f_bench.err
timeout 200 python bench.py --config cfg4g --steps 10 --warmup 3 --graph 1 --no-cpu-baseline > $O/r02f_bench_cfg4g.json 2>> $O/r02f_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv --log-file $O/r02f_launches_cfg3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r02f_ncu.log 2>&1
ls $O | grep r02f
