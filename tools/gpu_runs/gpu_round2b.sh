#!/bin/bash
mkdir -p gpurun_out
KT_GRAPH=0 timeout 300 python tools/bench_simple.py --case karman --nx 640 --ny 128 --steps 5 2>&1 | tail -1 | tee gpurun_out/karman640_direct.json
KT_GRAPH=1 timeout 300 python tools/bench_simple.py --case karman --nx 640 --ny 128 --steps 5 2>&1 | tail -1 | tee gpurun_out/karman640_graph.json
KT_GRAPH=0 timeout 300 python tools/bench_simple.py --case cavity --n 256 --steps 10 --alpha-p 0.005 2>&1 | tail -1 | tee gpurun_out/cavity256_direct.json
KT_GRAPH=1 timeout 300 python tools/bench_simple.py --case cavity --n 256 --steps 10 --alpha-p 0.005 2>&1 | tail -1 | tee gpurun_out/cavity256_graph.json
timeout 600 python bench.py --workload pcor4096 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
