#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/sweep10.txt; : > $S
for n in 16384 4096; do
  it=20; [ $n = 4096 ] && it=50
  timeout 200 python tools/microbench.py --n $n --iters $it --brief store_s >> $S 2>&1
  timeout 200 python tools/microbench.py --n $n --iters $it --no-s --brief no_s >> $S 2>&1
done
cut -c1-120 $S
timeout 600 python -m pytest tests/test_gpu_simple.py tests/test_gpu_krylov.py -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
