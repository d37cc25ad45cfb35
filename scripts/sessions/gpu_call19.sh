#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call19.log
: > $L
timeout 900 python -m pytest tests/test_gpu_flux2.py tests/test_gpu_flux.py tests/test_gpu_vae.py -m gpu -q --durations=5 >> $L 2>&1; echo "rc=$?" >> $L
timeout 200 python scripts/bench_flux.py --steps 10 --warmup 3 --graph > gpurun_out/bench_flux_b.json 2>> $L; echo "rc=$?" >> $L
cat gpurun_out/bench_flux_b.json >> $L
tail -c 4000 $L
