#!/bin/bash
# Runs every bring-up check in its own process under `timeout`; results -> gpurun_out/check.jsonl
mkdir -p gpurun_out
OUT=gpurun_out/check.jsonl
: > $OUT
run() { echo "== $*" >> $OUT; timeout 150 python scripts/gpu_check.py "$@" >> $OUT 2>gpurun_out/last_err.log || { echo "{\"test\": \"$*\", \"rc\": $?}" >> $OUT; tail -5 gpurun_out/last_err.log >> $OUT; }; }
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv >> $OUT
run elementwise
run linear 128 256 64 0
run linear 128 256 256 0
run linear 256 512 256 0
run linear 300 520 264 0
run linear 1024 1024 1024 1
run linear 1024 1024 1024 2
run linear 1000 64 512 3
run linear 8192 5120 5120 0 1
run linear 75600 5120 5120 0 1
run linear 75600 13824 5120 0 1
run attn 1 1 128 128
run attn 1 2 256 256
run attn 1 2 300 333
run attn 1 32 1024 1024 0 42
run attn 2 4 2048 512
run attn 1 4 8192 8192 1
run attn 1 40 75600 512 1
run attn 1 2 75600 75600 1
cat $OUT
