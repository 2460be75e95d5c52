#!/bin/bash
# One GPU-box pass that collects the round's evidence: gpu tests, the bench line, the ncu launch list of the bench command and
# `ncu --set full` captures of the dominant kernels (each only after its command has exited 0 without ncu).
# usage: tools/gpu_evidence.sh <tag>     -> gpurun_out/<tag>_*
tag=${1:-ev}
o=gpurun_out
python -m pytest tests -m gpu -x -q > $o/${tag}_gputest.log 2>&1; echo "pytest rc=$?" | tee -a $o/${tag}_gputest.log
python bench.py > $o/${tag}_bench1.json 2> $o/${tag}_bench1.err; echo "bench rc=$?"
BCMD="python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --e2e-frames 8 --bands-steps 2"
$BCMD > $o/${tag}_bench_short.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/${tag}_launches.csv $BCMD > $o/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sobel_tma_kernel -c 1 -s 3 -o $o/${tag}_bench_fused -f $BCMD --no-extras --no-bands > $o/${tag}_ncu_fused.log 2>&1
summ() {  # the .ncu-rep files are large (gpurun_out/ travels back only below 64 MiB): keep the text summaries
  python tools/ncu_summary.py $o/${tag}_$1.ncu-rep > $o/${tag}_$1.txt 2>&1
  python tools/ncu_hot.py $o/${tag}_$1.ncu-rep 12 2 > $o/${tag}_$1.hot.txt 2>&1
  rm -f $o/${tag}_$1.ncu-rep
}
summ bench_fused
cap() {  # name, kernel regex, prof.py args
  python tools/prof.py $3 > $o/${tag}_prof_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -s 3 -o $o/${tag}_$1 -f python tools/prof.py $3 --iters 2 > $o/${tag}_ncu_$1.log 2>&1
  summ $1
  tail -1 $o/${tag}_prof_$1.log
}
cap conv5 conv_tma "conv5"
cap conv5g conv_tma "conv5g"
cap conv9 conv_tma "conv9"
cap conv9g conv_tma "conv9g"
cap sobel2000 sobel_tma "sobel --width 2000 --height 3000 --frames 128"
cap sobel4k sobel_tma "sobel --width 3840 --height 2160 --frames 64"
cap pass sobel_tma "pass"
cap passes sobel_tma "passes"
cap gray stream_kernel "gray"
tail -3 $o/${tag}_gputest.log
