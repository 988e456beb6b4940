#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_errors.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
S=gpurun_out/sweep5.txt; : > $S
for n in 16384 4096 1024; do
  it=20; [ $n = 4096 ] && it=50; [ $n = 1024 ] && it=200
  timeout 200 python tools/microbench.py --n $n --iters $it --brief unfused >> $S 2>&1
  timeout 200 python tools/microbench.py --n $n --iters $it --fused --brief fused >> $S 2>&1
done
cat $S
timeout 900 python bench.py --steps 3 > gpurun_out/bench_32768_fused.json 2> gpurun_out/bench_32768_fused.err
tail -c 500 gpurun_out/bench_32768_fused.err; cut -c1-300 gpurun_out/bench_32768_fused.json
CMD="python bench.py --workload pcor16384 --steps 2 --warmup 3 --iters 10 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 50 -c 90 --csv \
  --log-file gpurun_out/r01_launches_bench16384_fused.csv $CMD > gpurun_out/ncu_launch.log 2>&1
