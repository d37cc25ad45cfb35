#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_call9.log
: > $L
run() { echo "== $1 :: ${@:2}" >> $L; env $1 timeout 120 "${@:2}" >> $L 2>&1 || echo "FAILED/TIMEOUT rc=$?" >> $L; }
run "B200_ATTN_PIPE=1" python scripts/attn_variant_ab.py bench
run "B200_ATTN_PIPE=1 APEX_B200_LIB=$PWD/apex-studio_b200/libapex_b200_p2.so" python scripts/attn_variant_ab.py bench
run "B200_ATTN_PIPE=1 APEX_B200_LIB=$PWD/apex-studio_b200/libapex_b200_p6.so" python scripts/attn_variant_ab.py bench
run "B200_ATTN_PIPE=0" python scripts/attn_variant_ab.py bench
run "B200_ATTN_PIPE=1" python scripts/attn_timeline.py 40 75600
grep -E "^==|FAIL" $L; grep -oE "\"self40\": \{[^}]*\}" $L; tail -1 $L
