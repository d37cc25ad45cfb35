#!/bin/bash
# Flux path: GPU test-suite (all families), FLUX.1-dev full-size step timing (eager + CUDA graph).
mkdir -p gpurun_out
L=gpurun_out/call9.log
: > $L
echo "== pytest gpu" >> $L
timeout 900 python -m pytest tests -m gpu -q --durations=8 >> $L 2>&1; echo "rc=$?" >> $L
echo "== bench_flux" >> $L
timeout 300 python scripts/bench_flux.py --steps 10 --warmup 3 --graph > gpurun_out/bench_flux.json 2>> $L; echo "rc=$?" >> $L
cat gpurun_out/bench_flux.json >> $L
tail -c 7000 $L
