#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/p2p_debug.py 2>&1 | grep -E "^rank|Error|error" | head -20 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log | cut -c1-250
for comm in p2p nccl; do
timeout 600 python bench.py --gpus 2 --steps 3 --workload pcor16384 --comm $comm --no-e2e 2> gpurun_out/bench_n2_16384_$comm.err | cut -c1-200
timeout 600 python bench.py --gpus 2 --steps 3 --workload pcor4096 --comm $comm --no-e2e 2>> gpurun_out/bench_n2_16384_$comm.err | cut -c1-200
done
grep -E "Error|error" gpurun_out/bench_n2_16384_p2p.err | head -5 | cut -c1-300
