#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_errors.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-250
S=gpurun_out/sweep9.txt; : > $S
for n in 48 64 128 176; do
  timeout 200 python tools/microbench.py --n $n --iters 60 --brief cluster16 >> $S 2>&1
  KT_CLUSTER_MAX=8 timeout 200 python tools/microbench.py --n $n --iters 60 --brief cluster8 >> $S 2>&1
done
cut -c1-110 $S
B=gpurun_out/simple_rates5.txt; : > $B
timeout 300 python tools/bench_simple.py --case cavity --n 50 --steps 20 >> $B 2>&1
timeout 300 python tools/bench_simple.py --case cavity --n 128 --alpha-p 0.02 --steps 10 >> $B 2>&1
timeout 300 python tools/bench_simple.py --case karman --nx 320 --ny 64 --steps 5 >> $B 2>&1
grep -v BICG $B | cut -c1-330
