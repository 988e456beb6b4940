#!/bin/bash
# final single-GPU check of round 1: full GPU suite, smoke, resident-solver latencies, default bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_final.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for n in 48 128 256 512; do timeout 120 python tools/microbench.py --n $n --iters 120 --no-s --brief "final n=$n" | cut -c1-110; done 2>&1 | tee gpurun_out/final_latency.txt
timeout 600 python bench.py --workload cavity128 --steps 20 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cavity128', d['value'])"
timeout 600 python bench.py --workload karman320 --steps 20 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('karman320', d['value'])"
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err; cut -c1-200 gpurun_out/bench_final.json
