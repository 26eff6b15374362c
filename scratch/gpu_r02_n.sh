#!/bin/bash
# round 2, GPU call N: ncu --set full of the conv kernels on the layers the review named (mid-resolution slab layers, kd-split layer, wgrad)
mkdir -p gpurun_out; O=gpurun_out
timeout 500 ncu --set full --clock-control none -k regex:'conv_tc' -c 40 -o $O/r02n_conv_layers python scratch/bench_conv.py d1.c1,d1.c2,u1.c2,u0.c1 1 > $O/r02n_ncu.log 2>&1
ncu -i $O/r02n_conv_layers.ncu-rep --page raw --csv > $O/r02n_conv_layers_raw.csv 2>/dev/null
rm -f $O/r02n_conv_layers.ncu-rep
ls -la $O | grep r02n
