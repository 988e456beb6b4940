#!/bin/bash
# final 8-GPU check with the round-1 end state: default bench (--comm p2p) at N = 8 and 4, reference arm under torchrun, CLI
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29811 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_final_n8.json 2> gpurun_out/bench_final_n8.err; echo "n8 rc=$?"
timeout 300 $TR --nproc-per-node 4 --master-port 29812 bench.py --gpus 4 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_final_n4.json 2> gpurun_out/bench_final_n4.err; echo "n4 rc=$?"
timeout 200 $TR --nproc-per-node 8 --master-port 29813 bench.py --gpus 8 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref_n8.json 2>/dev/null; echo "ref rc=$?"
for f in n8 n4 ref_n8; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_final_$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value'],1), (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('cores'), (d.get('config') or {}).get('parallelism'))
except Exception as e:
    print('$f FAILED', e); print(open('gpurun_out/bench_final_$f.err').read()[-1200:] if '$f' != 'ref_n8' else '')
PY
done
timeout 200 python -m karman_taichi_b200 karman --nx 640 --ny 128 --time-steps 1 --max-sub 5 2>&1 | tail -2
