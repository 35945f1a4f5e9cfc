#!/bin/bash
# First GPU bring-up: staged so that a trap in a tensor-core kernel cannot mask the fp32 results.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/env.txt 2>&1; nproc >> gpurun_out/env.txt; ls /root/reference >> gpurun_out/env.txt 2>&1
python __graft_entry__.py >> gpurun_out/env.txt 2>&1
echo "=== kernels (non-TC)"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "not tensor_core and not bfloat16 and not float16" 2>&1 | tail -25
echo "=== kernels (TC)"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tensor_core or bfloat16 or float16" 2>&1 | tail -40
echo "=== model fp32"; timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -s -k "fp32 or float_t or cfg0 or ema or bad_labels" 2>&1 | tail -40
echo "=== model 16-bit"; timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -s -k "bf16 or f16 or philox" 2>&1 | tail -40
