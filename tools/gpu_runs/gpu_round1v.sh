#!/bin/bash
# x-slab CG tests (2 GPUs) + SIMPLE workloads of bench.py
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab.py -x -q -m gpu > gpurun_out/pytest_v.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_v.log
for w in cavity128 karman320 karman16384; do
  st=20; [ $w = karman16384 ] && st=3
  timeout 600 python bench.py --workload $w --steps $st --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
  tail -c 1500 gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err
done
timeout 300 python bench.py --workload karman320 --impl reference --steps 3 --warmup 1 > gpurun_out/bench_karman320_ref.json 2>&1; tail -c 600 gpurun_out/bench_karman320_ref.json
