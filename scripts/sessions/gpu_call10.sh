#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call10.log
: > $L
echo "== pytest gpu hy15 + lora" >> $L
timeout 600 python -m pytest tests/test_gpu_hy15.py tests/test_lora.py -m gpu -q --durations=5 >> $L 2>&1; echo "rc=$?" >> $L
echo "== bench_hy15" >> $L
timeout 400 python scripts/bench_hy15.py --steps 2 --warmup 1 > gpurun_out/bench_hy15.json 2>> $L; echo "rc=$?" >> $L
cat gpurun_out/bench_hy15.json >> $L
tail -c 6000 $L
