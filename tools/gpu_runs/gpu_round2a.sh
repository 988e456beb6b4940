#!/bin/bash
# CUDA-graph replay of iteration chunks on mid-size grids
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_krylov.py tests/test_gpu_simple.py tests/test_gpu_karman.py tests/test_gpu_slab.py -m gpu -q -x > gpurun_out/pytest_2a.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_2a.log | cut -c1-300
for n in 200 256 512 1024 2048; do
  KT_GRAPH=0 timeout 120 python tools/microbench.py --n $n --iters 200 --no-s --brief "direct n=$n" | cut -c1-120
  KT_GRAPH=1 timeout 120 python tools/microbench.py --n $n --iters 200 --no-s --brief "graph  n=$n" | cut -c1-120
done 2>&1 | tee gpurun_out/graph_latency.txt
KT_GRAPH=0 timeout 300 python tools/bench_simple.py --case cavity --n 256 --steps 10 --alpha-p 0.02 2>&1 | tail -1 | tee gpurun_out/cavity256_direct.json
KT_GRAPH=1 timeout 300 python tools/bench_simple.py --case cavity --n 256 --steps 10 --alpha-p 0.02 2>&1 | tail -1 | tee gpurun_out/cavity256_graph.json
