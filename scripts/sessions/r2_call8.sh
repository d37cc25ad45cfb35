#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_call8.log
: > $L
run() { echo "== $1 :: ${@:2}" >> $L; env $1 timeout 120 "${@:2}" >> $L 2>&1 || echo "FAILED/TIMEOUT rc=$?" >> $L; }
run "B200_ATTN_PIPE=1" python scripts/attn_variant_ab.py
run "B200_ATTN_PIPE=1" python scripts/attn_variant_ab.py bench
run "B200_ATTN_PIPE=1 B200_ATTN_EXP=2" python scripts/attn_variant_ab.py bench
run "B200_ATTN_PIPE=0" python scripts/attn_variant_ab.py bench
run "B200_ATTN_PIPE=1 B200_ATTN_EXP=256 ATTN_TIMELINE_RAW=1" python scripts/attn_timeline.py 40 75600
cat $L
