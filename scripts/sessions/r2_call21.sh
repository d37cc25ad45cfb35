#!/bin/bash
# persistent attention grid: correctness (ragged, > 148 items), then A/B persistent vs one-CTA-per-item on the same box
mkdir -p gpurun_out
L=gpurun_out/r2_attn_persist_ab.log
: > $L
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "attention" 2>&1 | tail -3 | tee -a $L
for P in 1 0 1 0; do
  echo "== B200_ATTN_PERSIST=$P :: python scripts/attn_variant_ab.py bench" | tee -a $L
  B200_ATTN_PERSIST=$P timeout 300 python scripts/attn_variant_ab.py bench 2>&1 | tail -1 | tee -a $L
done
