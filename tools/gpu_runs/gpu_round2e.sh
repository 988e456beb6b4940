#!/bin/bash
# 2-GPU slab tests with the final library + ncu captures of the two resident solvers
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab.py -x -q -m gpu > gpurun_out/pytest_2e.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_2e.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cluster_solve -c 1 -o gpurun_out/r01_cluster_full \
  python tools/microbench.py --n 128 --iters 200 --brief x > gpurun_out/ncu_cluster.log 2>&1; echo "ncu cluster rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_coop_solve -c 1 -o gpurun_out/r01_coop_full \
  python tools/microbench.py --n 256 --iters 200 --brief x > gpurun_out/ncu_coop.log 2>&1; echo "ncu coop rc=$?"
ls -la gpurun_out/*.ncu-rep
