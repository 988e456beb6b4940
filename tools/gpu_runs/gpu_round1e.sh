#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
S=gpurun_out/sweep4.txt; : > $S
for n in 16384 4096 1024; do
  it=20; [ $n = 4096 ] && it=50; [ $n = 1024 ] && it=200
  for bb in 262144 524288 1048576 2097152 4194304; do
    KT_BLOCK_BYTES=$bb timeout 200 python tools/microbench.py --n $n --iters $it --brief bb$bb >> $S 2>&1
  done
done
timeout 600 python bench.py --workload pcor4096 --steps 3 > gpurun_out/bench_4096.json 2> gpurun_out/bench_4096.err
timeout 600 python bench.py --workload pcor16384 --steps 3 > gpurun_out/bench_16384.json 2> gpurun_out/bench_16384.err
timeout 900 python bench.py --steps 3 > gpurun_out/bench_32768.json 2> gpurun_out/bench_32768.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -5 gpurun_out/pytest_gpu.log
cat $S
tail -c 1500 gpurun_out/bench_4096.err; cut -c1-1500 gpurun_out/bench_4096.json
tail -c 600 gpurun_out/bench_32768.err; cut -c1-2500 gpurun_out/bench_32768.json
cut -c1-600 gpurun_out/bench_ref.json
