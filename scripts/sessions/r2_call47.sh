#!/bin/bash
# warp-uniform MMA issue loop (default build) vs the elected-lane loop (libapex_b200_elected.so), same box
mkdir -p gpurun_out
L=gpurun_out/r2_attn_uniform_issue_ab.log
: > $L
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "attention or scatter" 2>&1 | tail -2 | tee -a $L
run() { echo "== $* :: attn_variant_ab.py bench" | tee -a $L; env "$@" timeout 300 python scripts/attn_variant_ab.py bench 2>&1 | tail -1 | sed -E 's/"rel_[^,]*, //g; s/"nan": false, //' | tee -a $L; }
run APEX_B200_LIB=$PWD/apex-studio_b200/libapex_b200_elected.so
run X=uniform
run APEX_B200_LIB=$PWD/apex-studio_b200/libapex_b200_elected.so
run X=uniform
