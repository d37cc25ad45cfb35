#!/bin/bash
# Validation of the restored tree: smoke(), the whole GPU test-suite, the default N=1 bench.
mkdir -p gpurun_out
L=gpurun_out/call8.log
: > $L
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv >> $L
echo "== smoke" >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1; echo "rc=$?" >> $L
echo "== pytest gpu" >> $L
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 >> $L 2>&1; echo "rc=$?" >> $L
echo "== bench" >> $L
timeout 600 python bench.py > gpurun_out/bench_n1_call8.json 2>> $L; echo "rc=$?" >> $L
cat gpurun_out/bench_n1_call8.json >> $L
tail -c 6000 $L
