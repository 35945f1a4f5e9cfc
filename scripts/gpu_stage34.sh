#!/bin/bash
echo "=== vae tests"; timeout 900 python -m pytest tests/test_gpu_vae.py -m gpu -q -x -s 2>&1 | grep -E "passed|failed|Error|error|vae decode|DiffusionVAE|assert" | tail -30
echo "=== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -3
