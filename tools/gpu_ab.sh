#!/bin/bash
# A/B of the Sobel kernels: the round-1 library (lib_alt_r01, built from the r01 sources) against the current one.
# usage: tools/gpu_ab.sh [tag]      (run on the GPU box; writes gpurun_out/<tag>_ab.log)
tag=${1:-ab}
R01=openmp-parallel-computing_b200/lib_alt_r01/libompimg_b200.so
{
for k in sobel fused; do
  for shape in "3840 2160 64" "2000 3000 96" "1920 1080 256"; do
    set -- $shape
    [ -f $R01 ] && { echo -n "$k $1x$2x$3 r01: "; OMPIMG_LIB=$R01 python tools/prof.py $k --width $1 --height $2 --frames $3 --iters 20 2>&1 | tail -1; }
    echo -n "$k $1x$2x$3 new: "; python tools/prof.py $k --width $1 --height $2 --frames $3 --iters 20 2>&1 | tail -1
  done
done
} 2>&1 | tee gpurun_out/${tag}_ab.log
