#!/bin/bash
# final 2-GPU check: slab tests + bench at N = 2 (default --comm p2p) + reference arm under torchrun
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab.py -x -q -m gpu > gpurun_out/pytest_final2.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_final2.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29711 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_final_n2.json 2> gpurun_out/bench_final_n2.err; echo "n2 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_final_n2.json').read().strip().splitlines()[-1]); print('N=2', d['value'], d['e2e']['value'], d['config']['parallelism'], d['roofline']['frac'])"
timeout 300 $TR --nproc-per-node 2 --master-port 29712 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-200
