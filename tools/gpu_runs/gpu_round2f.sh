#!/bin/bash
# 256-thread register-resident cluster kernel (3..6 elements per thread)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_krylov.py tests/test_gpu_simple.py tests/test_gpu_karman.py -m gpu -q -x > gpurun_out/pytest_2f.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_2f.log | cut -c1-300
for n in 80 100 128 140 150; do
  timeout 120 python tools/microbench.py --n $n --iters 60 --brief "reg256 n=$n" | cut -c1-110
  KT_CLUSTER_REG=1 timeout 120 python tools/microbench.py --n $n --iters 60 --brief "smem   n=$n" | cut -c1-110
done 2>&1 | tee gpurun_out/cluster_reg256_latency.txt
for w in cavity128 karman320; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_r256.json 2> gpurun_out/bench_${w}_r256.err; echo "$w rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/bench_${w}_r256.json').read().strip().splitlines()[-1]); print('$w', d['value'], d['e2e']['value'], d['roofline']['krylov_iterations_per_step'])"
done
