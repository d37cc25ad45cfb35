#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_call11.log
: > $L
run() { echo "== $1 :: ${@:2}" >> $L; env $1 timeout 120 "${@:2}" >> $L 2>&1 || echo "FAILED/TIMEOUT rc=$?" >> $L; }
run "B200_ATTN_PIPE=1" python scripts/attn_variant_ab.py bench
for v in h1p0 h0p0 h0p1; do
run "B200_ATTN_PIPE=1 APEX_B200_LIB=$PWD/apex-studio_b200/libapex_b200_$v.so" python scripts/attn_variant_ab.py bench
done
run "B200_ATTN_PIPE=1" python scripts/attn_variant_ab.py bench
run "B200_ATTN_PIPE=1" python scripts/attn_timeline.py 40 75600
grep -E "^==|FAIL" $L; grep -oE "rel_1x32x1024x1024\": [0-9.]*|\"self40\": \{[^}]*\}" $L; tail -1 $L
