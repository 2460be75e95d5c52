# Tuning sweep of the producer-warp convolution kernel (conv_sep.cuh: dense-pitch layouts; OMPIMG_CONV_TMA=0 keeps aligned
# layouts on it too).  tools/conv_sweep2.sh sweeps the tensor-map kernel.
export OMPIMG_CONV_TMA=0
for k in conv5 conv9 conv5g conv9g; do
 for st in 4 6 8; do for rl in 1 0; do
  echo -n "stages=$st relax=$rl "; OMPIMG_CONV_STAGES=$st OMPIMG_CONV_RELAX=$rl python tools/prof.py $k --iters 20 2>&1 | tail -1
 done; done
 for band in 38 56 112 148; do
  echo -n "band=$band "; OMPIMG_CONV_BAND=$band python tools/prof.py $k --iters 20 2>&1 | tail -1
 done
done
