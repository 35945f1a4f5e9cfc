#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "groupnorm or gn or conv" 2>&1 | tail -3
timeout 300 python scripts/prof_layers.py 2>&1 | grep -E "total|gn_apply|conv3x3" | head -70
