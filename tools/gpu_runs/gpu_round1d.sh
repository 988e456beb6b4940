#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/sweep3.txt; : > $S
for v in w1_minb4 w1_minb5 w1_minb6 w2; do
  for c in 8 32 64 128 256 512; do
    KT_GRID_CAP=$c KT_LIB_PATH=$PWD/gpurun_variants/libkt_$v.so timeout 200 python tools/microbench.py --n 16384 --brief ${v}_cap$c >> $S 2>&1
  done
done
for ti in 8 16 32 128; do
  KT_TI=$ti KT_GRID_CAP=128 KT_LIB_PATH=$PWD/gpurun_variants/libkt_w1_minb4.so timeout 200 python tools/microbench.py --n 16384 --brief w1_minb4_cap128_TI$ti >> $S 2>&1
done
for c in 8 64 256; do
  KT_GRID_CAP=$c KT_LIB_PATH=$PWD/gpurun_variants/libkt_w1_minb4.so timeout 200 python tools/microbench.py --n 4096 --iters 50 --brief w1_minb4_cap$c >> $S 2>&1
  KT_GRID_CAP=$c KT_LIB_PATH=$PWD/gpurun_variants/libkt_w1_minb4.so timeout 200 python tools/microbench.py --n 1024 --iters 200 --brief w1_minb4_cap$c >> $S 2>&1
done
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
cat $S
