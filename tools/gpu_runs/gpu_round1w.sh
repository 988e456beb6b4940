#!/bin/bash
# cluster solver with one barrier less per iteration: parity + latency
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_w.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_w.log | cut -c1-300
for n in 48 64 128 176; do timeout 120 python tools/microbench.py --n $n --iters 60 --brief "cluster3 n=$n"; done 2>&1 | tee gpurun_out/cluster3_latency.txt
for w in cavity128 karman320; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_c3.json 2> gpurun_out/bench_${w}_c3.err; echo "$w rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/bench_${w}_c3.json').read().strip().splitlines()[-1]); print('$w', d['value'], d['e2e']['value'], d['roofline']['krylov_iterations_per_step'])"
done
python tools/bench_simple.py --case cavity --n 50 --steps 40 2>&1 | tail -1
