#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
timeout 600 python bench.py --gpus 8 --steps 3 > gpurun_out/bench_n8_32768.json 2> gpurun_out/bench_n8_32768.err
tail -c 400 gpurun_out/bench_n8_32768.err; cut -c1-250 gpurun_out/bench_n8_32768.json
timeout 600 python bench.py --gpus 4 --steps 3 > gpurun_out/bench_n4_32768.json 2> gpurun_out/bench_n4_32768.err
tail -c 400 gpurun_out/bench_n4_32768.err; cut -c1-250 gpurun_out/bench_n4_32768.json
