#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/sweep7.txt; : > $S
for n in 48 64 128 176; do
  KT_CLUSTER=1 timeout 200 python tools/microbench.py --n $n --iters 60 --brief cluster >> $S 2>&1
  KT_CLUSTER=0 timeout 200 python tools/microbench.py --n $n --iters 60 --brief stream >> $S 2>&1
done
cut -c1-150 $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cluster -c 2 -o gpurun_out/r01_cluster_full \
  python tools/microbench.py --n 128 --iters 60 --brief ncu > gpurun_out/ncu_cluster.log 2>&1
tail -3 gpurun_out/ncu_cluster.log
timeout 600 python -m pytest tests/test_gpu_simple.py -m gpu -q -x -k "checkpoint or first_steps" 2>&1 | tail -5
