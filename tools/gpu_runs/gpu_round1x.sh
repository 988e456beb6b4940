#!/bin/bash
# cluster solver with st.async + mbarrier exchange: parity, latency, segment profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_krylov.py tests/test_gpu_simple.py tests/test_gpu_karman.py -m gpu -q -x > gpurun_out/pytest_x.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_x.log | cut -c1-300
for n in 48 64 128 176; do timeout 120 python tools/microbench.py --n $n --iters 60 --brief "stas n=$n"; done 2>&1 | tee gpurun_out/cluster_stas_latency.txt
timeout 300 python tools/cluster_prof.py 48 128 176 2>&1 | tee gpurun_out/cluster_prof_stas.txt
for w in cavity128 karman320; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_stas.json 2> gpurun_out/bench_${w}_stas.err; echo "$w rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/bench_${w}_stas.json').read().strip().splitlines()[-1]); print('$w', d['value'], d['e2e']['value'], d['roofline']['krylov_iterations_per_step'])"
done
timeout 300 python tools/bench_simple.py --case cavity --n 50 --steps 40 2>&1 | tail -1
