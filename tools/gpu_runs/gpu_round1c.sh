#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/sweep2.txt; : > $S
for v in w1 w1_minb4 w1_su2 w2 w2_minb2 w2_minb3 w2_nv; do
  for w in 1 2 4 8; do
    KT_GRID_WAVES=$w KT_LIB_PATH=$PWD/gpurun_variants/libkt_$v.so timeout 200 python tools/microbench.py --n 16384 --brief ${v}_waves$w >> $S 2>&1
  done
done
for v in w1 w2 w2_minb2; do
  for w in 1 4; do
    KT_GRID_WAVES=$w KT_LIB_PATH=$PWD/gpurun_variants/libkt_$v.so timeout 200 python tools/microbench.py --n 4096 --iters 50 --brief ${v}_waves$w >> $S 2>&1
  done
done
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
cat $S
