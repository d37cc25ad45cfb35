#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call12.log
: > $L
timeout 600 python -m pytest tests/test_gpu_qwen.py tests/test_gpu_hy15.py -m gpu -q --durations=5 >> $L 2>&1; echo "rc=$?" >> $L
echo "== bench_qwen" >> $L
timeout 300 python scripts/bench_qwen.py > gpurun_out/bench_qwen.json 2>> $L; echo "rc=$?" >> $L
cat gpurun_out/bench_qwen.json >> $L
tail -c 5000 $L
