#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call32.log
: > $L
B200_ATTN_VARIANT=7 timeout 25 python scripts/attn_variant_ab.py 2>&1 | tail -2 | cut -c1-600 >> $L; echo "rc=${PIPESTATUS[0]}" >> $L
if grep -q '"nan": false' $L; then
  B200_ATTN_VARIANT=7 timeout 40 python scripts/attn_variant_ab.py bench 2>&1 | tail -1 | cut -c1-700 >> $L
  B200_ATTN_VARIANT=2 timeout 40 python scripts/attn_variant_ab.py bench 2>&1 | tail -1 | cut -c1-700 >> $L
fi
cat $L
