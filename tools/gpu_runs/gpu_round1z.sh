#!/bin/bash
# cluster solver: register kernel (<= 2 elements / thread) + shared-memory kernel: parity, latency, rates
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_krylov.py tests/test_gpu_simple.py tests/test_gpu_karman.py -m gpu -q -x > gpurun_out/pytest_z.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_z.log | cut -c1-300
for n in 32 48 64 100 128 176; do timeout 120 python tools/microbench.py --n $n --iters 60 --brief "final n=$n"; done 2>&1 | tee gpurun_out/cluster_final_latency.txt
for n in 48 64 100; do KT_CLUSTER_REG=0 timeout 120 python tools/microbench.py --n $n --iters 60 --brief "smem-kernel n=$n"; done 2>&1 | tee -a gpurun_out/cluster_final_latency.txt
for w in cavity128 karman320; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_final.json 2> gpurun_out/bench_${w}_final.err; echo "$w rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/bench_${w}_final.json').read().strip().splitlines()[-1]); print('$w', d['value'], d['e2e']['value'], d['roofline']['krylov_iterations_per_step'])"
done
timeout 300 python tools/bench_simple.py --case cavity --n 50 --steps 40 2>&1 | tail -1
