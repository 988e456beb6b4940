#!/bin/bash
# profiles for the final round-1 configuration + SIMPLE outer-iteration rates
mkdir -p gpurun_out
CMD="python bench.py --workload pcor16384 --steps 2 --warmup 3 --iters 10 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 120 --csv \
  --log-file gpurun_out/r01_launches_bench16384.csv $CMD > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stencil -s 20 -c 2 -o gpurun_out/r01_stencil_full \
  $CMD > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bicg_xr -s 10 -c 1 -o gpurun_out/r01_xr_full \
  $CMD > gpurun_out/ncu_full2.log 2>&1
B=gpurun_out/simple_rates.txt; : > $B
timeout 300 python tools/bench_simple.py --case cavity --n 50 --steps 20 >> $B 2>&1
timeout 300 python tools/bench_simple.py --case cavity --n 128 --alpha-p 0.02 --steps 10 >> $B 2>&1
timeout 300 python tools/bench_simple.py --case karman --nx 320 --ny 64 --steps 5 >> $B 2>&1
timeout 600 python tools/bench_simple.py --case karman --nx 16384 --ny 3276 --steps 2 --warmup 1 --krylov-cap 200 >> $B 2>&1
cat $B
timeout 300 python tools/microbench.py --n 16384 --kernels > gpurun_out/mb_16384_final.json 2>&1
tail -5 gpurun_out/ncu_launch.log
ls -la gpurun_out | tail -12
