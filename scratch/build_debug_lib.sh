#!/bin/bash
# Ablation build of the library: -DICH_TC_DEBUG compiles the ICH_TC_DBG switches (1 = no MMA issue, 2 = no TMA loads, 4 = no epilogue
# stores) back into the tcgen05 conv kernels.  Use with ICH_B200_LIB=<this .so>; results are WRONG by construction -- timing experiments only.
set -e
cd "$(dirname "$0")/.."
PKG=label-efficient-volumetric-deep-semantic-segmentation-of-ich_b200
OUT=$PKG/ich_b200/libich_b200_dbg.so
mkdir -p $PKG/build_dbg
for f in api gemm_generic elementwise loss conv_tc conv_tc_stream conv_cin1_tc aux_ops; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DICH_TC_DEBUG -I include -c $PKG/csrc/$f.cu -o $PKG/build_dbg/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o $OUT $PKG/build_dbg/*.o -lcuda -lcudart
echo $OUT
