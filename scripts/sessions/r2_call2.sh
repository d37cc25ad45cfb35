#!/bin/bash
# round 2, call 2: CTA-pair attention bring-up -- sanity in a subprocess under timeout, then A/B vs the 1-CTA kernel and cuDNN
mkdir -p gpurun_out
L=gpurun_out/r2_call2.log
: > $L
for pair in 1 0; do
  echo "== B200_ATTN_2CTA=$pair sanity" >> $L
  B200_ATTN_2CTA=$pair timeout 120 python scripts/attn_variant_ab.py >> $L 2>&1 || echo "FAILED/TIMEOUT rc=$?" >> $L
done
for pair in 1 0 1; do
  echo "== B200_ATTN_2CTA=$pair bench" >> $L
  B200_ATTN_2CTA=$pair timeout 300 python scripts/attn_variant_ab.py bench >> $L 2>&1 || echo "FAILED/TIMEOUT rc=$?" >> $L
done
cat $L
