#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_errors.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log | cut -c1-250
B=gpurun_out/simple_rates2.txt; : > $B
for c in 1 0; do
echo "KT_CLUSTER=$c" >> $B
KT_CLUSTER=$c timeout 300 python tools/bench_simple.py --case cavity --n 50 --steps 20 >> $B 2>&1
KT_CLUSTER=$c timeout 300 python tools/bench_simple.py --case cavity --n 128 --alpha-p 0.02 --steps 10 >> $B 2>&1
KT_CLUSTER=$c timeout 300 python tools/bench_simple.py --case karman --nx 320 --ny 64 --steps 5 >> $B 2>&1
done
grep -v BICG $B | cut -c1-330
