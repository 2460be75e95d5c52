#!/bin/bash
# tools/dropin_bench.sh -- the reference's own CLI timing line ("Compute kernel xN: S s", main.c:39 /
# main_with_sobel.c:79) for the reference build and for the same unmodified driver linked against
# libompimg_b200.so.  Needs oracle/_ref/bin (built by `make -C oracle ref` where /root/reference exists).
set -u
cd "$(dirname "$0")/.."
BIN=oracle/_ref/bin; FIX=oracle/_ref/fixtures; OUT=${TMPDIR:-/tmp}/dropin_$$; mkdir -p "$OUT"
P1=${1:-200}; P2=${2:-20}
echo "# host: $(nproc) cores; passes: grayscale x$P1 (test.jpg 1920x1080), gray+sobel x$P2 (2000x3000)"
for t in 1 $(nproc); do
  echo "reference -O3 -fopenmp, OMP_NUM_THREADS=$t: $(OMP_NUM_THREADS=$t $BIN/grayscale_strict $FIX/test.jpg $OUT/a.png $P1 | grep Compute)"
  echo "reference sobel driver,  OMP_NUM_THREADS=$t: $(OMP_NUM_THREADS=$t $BIN/sobel_strict $FIX/more_than_one_mega_photo.jpg $OUT/b.png $P2 | grep Compute)"
done
echo "libompimg_b200 (pageable, default):        $($BIN/grayscale_b200 $FIX/test.jpg $OUT/c.png $P1 | grep Compute)"
echo "libompimg_b200 (OMPIMG_PIN=register):      $(OMPIMG_PIN=register $BIN/grayscale_b200 $FIX/test.jpg $OUT/d.png $P1 | grep Compute)"
echo "libompimg_b200 sobel driver (pageable):    $($BIN/sobel_b200 $FIX/more_than_one_mega_photo.jpg $OUT/e.png $P2 | grep Compute)"
echo "libompimg_b200 sobel driver (PIN=register):$(OMPIMG_PIN=register $BIN/sobel_b200 $FIX/more_than_one_mega_photo.jpg $OUT/f.png $P2 | grep Compute)"
md5sum $OUT/a.png $OUT/c.png $OUT/d.png | awk '{print $1}' | sort -u | wc -l | sed 's/^/distinct grayscale PNGs (1 = identical): /'
md5sum $OUT/b.png $OUT/e.png $OUT/f.png | awk '{print $1}' | sort -u | wc -l | sed 's/^/distinct sobel PNGs (1 = identical): /'
rm -rf "$OUT"
