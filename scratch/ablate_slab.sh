#!/bin/bash
for dbg in 0 1 2 4 7; do
  echo "== ICH_TC_DBG=$dbg"
  ICH_TC_DBG=$dbg timeout 120 python scratch/bench_conv.py d1.c1,d1.c2,u1.c1,u1.c2,u0.c1,bt.c2 2>&1 | grep -v total
done
