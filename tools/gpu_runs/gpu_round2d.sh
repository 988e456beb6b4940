#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_2d.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_2d.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for n in 512 600 660; do
  timeout 120 python tools/microbench.py --n $n --iters 200 --no-s --brief "coop   n=$n" | cut -c1-120
  KT_COOP=0 timeout 120 python tools/microbench.py --n $n --iters 200 --no-s --brief "graph  n=$n" | cut -c1-120
done 2>&1 | tee gpurun_out/coop_latency2.txt
timeout 900 python bench.py > gpurun_out/bench_default2.json 2> gpurun_out/bench_default2.err; tail -c 300 gpurun_out/bench_default2.err; cut -c1-260 gpurun_out/bench_default2.json
