#!/bin/bash
for st in 0 300 700 1500; do
  echo "=== stagger=$st"
  SGB200_ATTN_STAGGER_NS=$st timeout 120 python scripts/prof_kernels.py attention 256 2>&1 | tail -3
done
