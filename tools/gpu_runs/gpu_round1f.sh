#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_errors.txt
nvidia-smi -L > gpurun_out/gpus.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
cat gpurun_out/parity_errors.txt
timeout 900 python bench.py --gpus 2 --steps 3 --workload pcor16384 > gpurun_out/bench_n2_16384.json 2> gpurun_out/bench_n2_16384.err
tail -c 1200 gpurun_out/bench_n2_16384.err; cut -c1-400 gpurun_out/bench_n2_16384.json
timeout 900 python bench.py --gpus 2 --steps 3 > gpurun_out/bench_n2_32768.json 2> gpurun_out/bench_n2_32768.err
tail -c 600 gpurun_out/bench_n2_32768.err; cut -c1-400 gpurun_out/bench_n2_32768.json
timeout 900 python bench.py --gpus 1 --steps 3 --no-e2e > gpurun_out/bench_n1_32768.json 2> gpurun_out/bench_n1_32768.err
cut -c1-300 gpurun_out/bench_n1_32768.json
