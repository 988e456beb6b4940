#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_errors.txt
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log | cut -c1-250
cat gpurun_out/parity_errors.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 300 gpurun_out/bench_default.err; cut -c1-400 gpurun_out/bench_default.json
timeout 300 python bench.py --impl reference > gpurun_out/bench_default_ref.json 2> gpurun_out/bench_default_ref.err
cut -c1-200 gpurun_out/bench_default_ref.json
