#!/bin/bash
for dbg in 0 1 2; do
  echo "== ICH_TC_DBG=$dbg"
  ICH_TC_DBG=$dbg timeout 120 python scratch/bench_conv.py 2>&1 | grep -v total | sed 's/| fwd.*| wgrad/| wgrad/'
done
