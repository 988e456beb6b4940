#!/bin/bash
# First GPU visit: environment facts, parity tests, micro-benchmarks, ncu launch list + one full capture.
mkdir -p gpurun_out
{ nvidia-smi; free -g; nproc; lscpu | head -20; } > gpurun_out/env.txt 2>&1
python -c "import karman_taichi_b200.build as b; b.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/microbench.py --n 4096 --kernels > gpurun_out/mb_4096.json 2> gpurun_out/mb_4096.err
timeout 600 python tools/microbench.py --n 16384 --kernels > gpurun_out/mb_16384.json 2> gpurun_out/mb_16384.err
timeout 300 python tools/microbench.py --n 1024 --kernels > gpurun_out/mb_1024.json 2> gpurun_out/mb_1024.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv \
  --log-file gpurun_out/launches_16384.csv python tools/microbench.py --n 16384 --iters 4 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stencil -s 4 -c 3 -o gpurun_out/prof_stencil \
  python tools/microbench.py --n 16384 --iters 4 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
tail -5 gpurun_out/pytest_gpu.log
cat gpurun_out/mb_16384.json | head -60
