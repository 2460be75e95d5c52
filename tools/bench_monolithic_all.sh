#!/bin/bash
# tools/bench_monolithic_all.sh -- BASELINE config 1 (monolithic grayscale on images/test.jpg, OMP_NUM_THREADS 1/2/4/6) in the
# reference's own CSV schema, for: the reference build, the same main.c linked against libompimg_b200 (own CUDA context per
# process), and the latter behind a running ompimg_worker.  Needs oracle/_ref/bin (make -C oracle ref where /root/reference exists).
set -u
cd "$(dirname "$0")/.."
OUT=${1:-gpurun_out}; RUNS=${2:-5}; PASSES=${3:-1}
BIN=oracle/_ref/bin; IMG=oracle/_ref/fixtures/test.jpg; LIBD=openmp-parallel-computing_b200/lib
mkdir -p "$OUT"
python tools/bench_monolithic.py --exe $BIN/grayscale_strict --image $IMG --threads 1 2 4 6 --runs $RUNS --passes $PASSES --csv $OUT/monolithic_reference.csv | tail -5
python tools/bench_monolithic.py --exe $BIN/grayscale_b200 --image $IMG --threads 1 6 --runs $RUNS --passes $PASSES --csv $OUT/monolithic_b200_own_context.csv | tail -3
SOCK=$(mktemp -u /tmp/ompimg_XXXX.sock)
$LIBD/ompimg_worker --socket $SOCK 2>/dev/null & WPID=$!
for i in $(seq 1 600); do [ -S $SOCK ] && break; sleep 0.1; done
OMPIMG_WORKER=$SOCK python tools/bench_monolithic.py --exe $BIN/grayscale_b200 --image $IMG --threads 1 2 4 6 --runs $RUNS --passes $PASSES --csv $OUT/monolithic_b200_worker.csv | tail -5
kill $WPID; wait $WPID 2>/dev/null
