#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_call3.log
: > $L
for pair in 0 1; do
  for hs in "40 75600" "2 75600"; do
    echo "== timeline B200_ATTN_2CTA=$pair heads/S=$hs" >> $L
    B200_ATTN_2CTA=$pair timeout 120 python scripts/attn_timeline.py $hs >> $L 2>&1 || echo "FAILED rc=$?" >> $L
  done
done
cat $L
