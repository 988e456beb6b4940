#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab.py tests/test_gpu_krylov.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log | cut -c1-250
for comm in p2p nccl; do
timeout 600 python bench.py --gpus 2 --steps 3 --workload pcor16384 --comm $comm --no-e2e > gpurun_out/bench_n2_16384_$comm.json 2> gpurun_out/bench_n2_16384_$comm.err
tail -c 600 gpurun_out/bench_n2_16384_$comm.err; cut -c1-200 gpurun_out/bench_n2_16384_$comm.json
done
timeout 600 python bench.py --gpus 2 --steps 3 --workload pcor4096 --comm p2p --no-e2e | cut -c1-200
timeout 600 python bench.py --gpus 2 --steps 3 --workload pcor4096 --comm nccl --no-e2e | cut -c1-200
timeout 900 python bench.py --gpus 2 --steps 3 > gpurun_out/bench_n2_32768_p2p.json 2> gpurun_out/bench_n2_32768_p2p.err
tail -c 300 gpurun_out/bench_n2_32768_p2p.err; cut -c1-200 gpurun_out/bench_n2_32768_p2p.json
