#!/bin/bash
# Second GPU visit: full parity suite + tuning-variant sweep.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
S=gpurun_out/sweep.txt; : > $S
for v in base minb3 minb4 su1 su1_minb4 su4 ew2 ew1 ew8; do
  KT_LIB_PATH=$PWD/gpurun_variants/libkt_$v.so timeout 200 python tools/microbench.py --n 16384 --brief $v >> $S 2>&1
done
for ti in 8 16 32 128 256; do
  KT_TI=$ti KT_LIB_PATH=$PWD/gpurun_variants/libkt_base.so timeout 200 python tools/microbench.py --n 16384 --brief base_TI$ti >> $S 2>&1
done
for w in 2 4; do
  KT_GRID_WAVES=$w KT_LIB_PATH=$PWD/gpurun_variants/libkt_base.so timeout 200 python tools/microbench.py --n 16384 --brief base_waves$w >> $S 2>&1
done
for v in base minb4 su1; do
  KT_LIB_PATH=$PWD/gpurun_variants/libkt_$v.so timeout 200 python tools/microbench.py --n 4096 --iters 50 --brief $v >> $S 2>&1
done
KT_LIB_PATH=$PWD/gpurun_variants/libkt_base.so timeout 200 python tools/microbench.py --n 16384 --kernels > gpurun_out/mb_16384_base.json 2>&1
tail -4 gpurun_out/pytest_gpu.log
cat $S
