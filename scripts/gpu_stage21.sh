#!/bin/bash
for cfg in "8 2 0" "8 2 1" "8 2 3" "8 0 3" "8 3 3"; do
  set -- $cfg
  echo "=== microbench attention v$1 poly8=$2 dbg=$3"
  SGB200_ATTN=$1 SGB200_ATTN_POLY8=$2 SGB200_ATTN_DBG=$3 timeout 300 python scripts/prof_kernels.py attention 128 2>&1 | tail -3
done
