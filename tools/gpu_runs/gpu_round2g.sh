#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_karman.py tests/test_gpu_krylov.py -m gpu -q -x -k "karman or jacobi or parity" > gpurun_out/pytest_2g.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_2g.log | cut -c1-300
for n in 150; do timeout 120 python tools/microbench.py --n $n --iters 60 --brief "lean6 n=$n" | cut -c1-110; done
timeout 600 python bench.py --workload karman320 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('karman320', d['value'], d['roofline']['krylov_iterations_per_step'])"
