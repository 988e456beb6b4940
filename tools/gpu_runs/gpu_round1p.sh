#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_slab.py -m gpu -q -k p2p_single 2>&1 | tail -15 | cut -c1-250
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_slab.py -m gpu -q -k p2p_single 2>&1 | grep -E "Invalid|Error|ERROR SUMMARY|at kt|k_p2p" | head -20 | cut -c1-250
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/p2p_debug.py 2>&1 | grep -E "^rank|Error|error" | head -30 | cut -c1-300
