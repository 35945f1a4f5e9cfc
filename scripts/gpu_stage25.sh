#!/bin/bash
mkdir -p gpurun_out
echo "=== fused kernel tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -x -k "fused or layernorm" 2>&1 | grep -E "passed|failed|Error|error|rel-L2|timed out|assert" | tail -30
echo "=== all kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q 2>&1 | tail -5
echo "=== mem microbench"; timeout 300 python scripts/prof_mem.py 1024 2>&1 | grep -E "mode1|layernorm|gelu=True"
