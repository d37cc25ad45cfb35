#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call26.log
: > $L
run() { echo "== $1 linear ${@:2}" >> $L; env $1 timeout 40 python scripts/gpu_check.py linear "${@:2}" 2>&1 | grep -v Warning | tail -2 | cut -c1-330 >> $L; echo "rc=${PIPESTATUS[0]}" >> $L; }
run "B200_LINEAR_2CTA=1" 300 520 264 0
run "B200_LINEAR_2CTA=1" 1024 1024 1024 2
run "B200_LINEAR_2CTA=1" 8192 5120 5120 0 1
run "B200_LINEAR_2CTA=1" 75600 5120 5120 0 1
tail -c 2500 $L
