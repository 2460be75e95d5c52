# old row-ring kernel (OMPIMG_SOBEL_WARP=0) against the warp-autonomous one (=1), with and without the folded last strip
for k in sobel fused; do
  echo -n "warp=0        "; OMPIMG_SOBEL_WARP=0 python tools/prof.py $k --iters 20 2>&1 | tail -1
  echo -n "warp=1 fold=0 "; OMPIMG_SOBEL_WARP=1 OMPIMG_SOBEL_FOLD=0 python tools/prof.py $k --iters 20 2>&1 | tail -1
  echo -n "warp=1 fold=1 "; OMPIMG_SOBEL_WARP=1 OMPIMG_SOBEL_FOLD=1 python tools/prof.py $k --iters 20 2>&1 | tail -1
done
