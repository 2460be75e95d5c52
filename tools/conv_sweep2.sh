#!/bin/bash
# Sweep of the conv_tma kernel's tuning knobs (ring depth in multiples of (K+1)/2 stages, band height) on cfg4; run on the GPU box.
# usage: tools/conv_sweep2.sh [tag]   -> gpurun_out/<tag>_convsweep.log
tag=${1:-s}
{
for k in conv5 conv5g conv9 conv9g; do
  for st in 1 2 3; do
    for b in 0 64 96 128 180 222 300 400; do
      echo -n "$k stages=$st band=$b: "
      OMPIMG_CONV_STAGES=$st OMPIMG_CONV_BAND=$b python tools/prof.py $k --iters 20 2>&1 | tail -1
    done
  done
done
} 2>&1 | tee gpurun_out/${tag}_convsweep.log
