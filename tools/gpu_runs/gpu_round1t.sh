#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_errors.txt
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 640 -c 404 --csv \
  --log-file gpurun_out/r01_launches_bench_default.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_default.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stencil -s 20 -c 2 -o gpurun_out/r01_stencil_full \
  python bench.py --workload pcor16384 --steps 2 --warmup 3 --iters 10 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 300 gpurun_out/bench_default.err; cut -c1-300 gpurun_out/bench_default.json
timeout 300 python -m karman_taichi_b200 cavity --n 32 --max-steps 3 --frame gpurun_out/cavity.png 2>&1 | tail -3
