#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call25.log
: > $L
run() { echo "== $1 linear ${@:2}" >> $L; env $1 B200_LINEAR_DEBUG=1 timeout 40 python scripts/gpu_check.py linear "${@:2}" 2>&1 | grep -v Warning | tail -3 | cut -c1-400 >> $L; }
run "B200_LINEAR_2CTA=1" 8192 5120 5120 0 1
run "B200_LINEAR_2CTA=1 B200_LINEAR_PAIRS=37" 8192 5120 5120 0 1
run "B200_LINEAR_2CTA=1 B200_LINEAR_PAIRS=64" 8192 5120 5120 0 1
run "B200_LINEAR_2CTA=1 B200_LINEAR_PAIRS=72" 8192 5120 5120 0 1
tail -c 3000 $L
