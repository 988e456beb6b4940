#!/bin/bash
# 8-GPU validation of the default (--comm p2p) bench path, N = 8 and 4, NCCL at N = 8 for comparison, multi-GPU tests
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
timeout 300 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_p2p_n8.json 2> gpurun_out/bench_p2p_n8.err; echo "n8 p2p rc=$?"
tail -c 600 gpurun_out/bench_p2p_n8.json
timeout 300 $TR --nproc-per-node 4 --master-port 29612 bench.py --gpus 4 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_p2p_n4.json 2> gpurun_out/bench_p2p_n4.err; echo "n4 p2p rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29613 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --comm nccl > gpurun_out/bench_nccl_n8.json 2> gpurun_out/bench_nccl_n8.err; echo "n8 nccl rc=$?"
timeout 300 python -m pytest tests/test_gpu_slab.py tests/test_gpu_krylov.py -x -q -m gpu -k "slab or stepwise or p2p or nccl" > gpurun_out/pytest_u.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_u.log
for f in p2p_n8 p2p_n4 nccl_n8; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value'],1), d['ms_per_step'], (d.get('e2e') or {}).get('value'), d['roofline']['frac'], d['config']['parallelism'])
except Exception as e:
    print('$f', 'FAILED', e); print(open('gpurun_out/bench_$f.err').read()[-1500:])
PY
done
