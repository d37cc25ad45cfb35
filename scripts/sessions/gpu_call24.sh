#!/bin/bash
# CTA-pair (cta_group::2) linear kernel bring-up: each check in its own process under a short timeout (a protocol error hangs)
mkdir -p gpurun_out
L=gpurun_out/call24.log
: > $L
run() { echo "== 2cta=$1 linear ${@:2}" >> $L; B200_LINEAR_2CTA=$1 timeout 40 python scripts/gpu_check.py linear "${@:2}" 2>&1 | tail -2 | cut -c1-700 >> $L; echo "rc=${PIPESTATUS[0]}" >> $L; }
run 1 256 256 64 0
if grep -q '"rel_l2": 0.00' $L; then
  run 1 256 512 256 0
  run 1 300 520 264 0
  run 1 1024 1024 1024 1
  run 1 1024 1024 1024 2
  run 1 1000 64 512 3
  run 1 130 256 128 0
  run 1 8192 5120 5120 0 1
  run 0 8192 5120 5120 0 1
  run 1 75600 5120 5120 0 1
  run 0 75600 5120 5120 0 1
  run 1 4096 3072 12288 2 1
  run 0 4096 3072 12288 2 1
fi
tail -c 6000 $L
