#!/bin/bash
# persistent cooperative solver for mid-size grids: parity + latency + SIMPLE rates
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_krylov.py tests/test_gpu_simple.py tests/test_gpu_karman.py -m gpu -q -x > gpurun_out/pytest_2c.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_2c.log | cut -c1-300
for n in 200 256 384 512; do
  timeout 120 python tools/microbench.py --n $n --iters 200 --no-s --brief "coop   n=$n" | cut -c1-120
  KT_COOP=0 timeout 120 python tools/microbench.py --n $n --iters 200 --no-s --brief "graph  n=$n" | cut -c1-120
done 2>&1 | tee gpurun_out/coop_latency.txt
timeout 300 python tools/bench_simple.py --case karman --nx 640 --ny 128 --steps 5 2>&1 | tail -1 | tee gpurun_out/karman640_coop.json
timeout 300 python tools/bench_simple.py --case cavity --n 256 --steps 10 --alpha-p 0.005 2>&1 | tail -1 | tee gpurun_out/cavity256_coop.json
