#!/bin/bash
echo "=== attention kernel tests (v1)"; SGB200_ATTN=1 timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention_tensor_core" 2>&1 | tail -2
for v in 1 3; do echo "=== microbench attention v$v"; SGB200_ATTN=$v SGB200_ATTN_POLY=0 python scripts/prof_kernels.py attention 128 2>&1 | tail -3; done
echo "=== v1 poly 4"; SGB200_ATTN=1 SGB200_ATTN_POLY=4 python scripts/prof_kernels.py attention 128 2>&1 | tail -3
