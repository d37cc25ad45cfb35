#!/bin/bash
# staged epilogue (SW64 quarter boxes) + persistent grid, zero-spill build, vs the round-2 HEAD kernel on the same box
mkdir -p gpurun_out
L=gpurun_out/r2_attn_stage_ab2.log
: > $L
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "attention" 2>&1 | tail -3 | tee -a $L
run() { echo "== $* :: attn_variant_ab.py bench" | tee -a $L; env "$@" timeout 300 python scripts/attn_variant_ab.py bench 2>&1 | tail -1 | tee -a $L; }
run APEX_B200_LIB=$PWD/apex-studio_b200/libapex_b200_head.so
run B200_ATTN_PERSIST=1
run B200_ATTN_PERSIST=0
run B200_ATTN_STAGE_KV=100
run APEX_B200_LIB=$PWD/apex-studio_b200/libapex_b200_head.so
run B200_ATTN_PERSIST=1
echo "== timeline persistent staged 40 x 75600 x 512" | tee -a $L
timeout 120 python scripts/attn_timeline_items.py 40 75600 512 20 2 2>&1 | tee -a $L
