#!/bin/bash
# Alternative code paths: the end-to-end GPU tests under every documented environment switch, and the attention kernel
# tests under every selectable attention version (last run: all green, 12 x 42 + 4 x 38 tests).
for cfg in "SGB200_ATTN=8" "SGB200_ATTN=11" "SGB200_ATTN=3" "SGB200_ATTN=1" "SGB200_PDL=1" "SGB200_PDL=0" "SGB200_FUSED_SA=0" "SGB200_FUSED_C=64" "SGB200_FUSED_OUTC=0" "SGB200_VCAT=0" "SGB200_RAW16=0" "SGB200_SHARED_PREFIX=0"; do
  echo "== $cfg: $(env $cfg timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_vae.py -m gpu -q -x 2>&1 | tail -1 | cut -c1-200)"
done
echo "== attention kernel tests under each version"
for v in 1 3 8 11; do echo "SGB200_ATTN=$v: $(SGB200_ATTN=$v timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k attention 2>&1 | tail -1 | cut -c1-200)"; done
