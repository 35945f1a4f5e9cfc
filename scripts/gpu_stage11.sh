#!/bin/bash
echo "=== attention kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention_tensor_core" 2>&1 | tail -2
for p in 0 4; do echo "=== microbench attention poly=$p"; SGB200_ATTN_POLY=$p python scripts/prof_kernels.py attention 128 2>&1 | tail -3; done
