#!/bin/bash
for dbg in 0 4; do
  echo "=== attention dbg=$dbg"
  SGB200_ATTN_DBG=$dbg timeout 120 python scripts/prof_kernels.py attention 128 2>&1 | tail -3
done
