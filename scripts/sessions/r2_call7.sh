#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_call7.log
: > $L
run() { echo "== $1 :: ${@:2}" >> $L; env $1 timeout 180 "${@:2}" >> $L 2>&1 || echo "FAILED/TIMEOUT rc=$?" >> $L; }
run "B200_ATTN_PIPE=1 B200_ATTN_2CTA=0" python scripts/attn_variant_ab.py
run "B200_ATTN_PIPE=1 B200_ATTN_2CTA=1" python scripts/attn_variant_ab.py
run "B200_ATTN_PIPE=1 B200_ATTN_2CTA=0" python scripts/attn_variant_ab.py bench
run "B200_ATTN_PIPE=1 B200_ATTN_2CTA=1" python scripts/attn_variant_ab.py bench
run "B200_ATTN_PIPE=0 B200_ATTN_2CTA=0" python scripts/attn_variant_ab.py bench
run "B200_ATTN_PIPE=1 B200_ATTN_2CTA=0" python scripts/attn_timeline.py 40 75600
run "B200_ATTN_PIPE=1 B200_ATTN_2CTA=1" python scripts/attn_timeline.py 40 75600
cat $L
