# A/B of builds of the library on the cfg4 convolutions (tools/prof.py), default build first
for k in conv5 conv5g conv9 conv9g; do
  echo -n "default "; python tools/prof.py $k --iters 20 2>&1 | tail -1
  for alt in lib_alt lib_alt2; do
    so=openmp-parallel-computing_b200/$alt/libompimg_b200.so
    if [ -f $so ]; then echo -n "$alt "; OMPIMG_LIB=$so python tools/prof.py $k --iters 20 2>&1 | tail -1; fi
  done
done
