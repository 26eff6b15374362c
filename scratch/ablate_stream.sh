#!/bin/bash
# ablation of the plane-streaming conv kernel: which of MMA issue / TMA / epilogue bounds it
for dbg in 0 1 2 4 3 5 6 7; do
  echo "== ICH_TC_DBG=$dbg"
  ICH_TC_DBG=$dbg timeout 120 python scratch/bench_conv.py d0.c2,u2.c1,u2.c2 2>&1 | grep -v total | sed 's/| wgrad.*//'
done
