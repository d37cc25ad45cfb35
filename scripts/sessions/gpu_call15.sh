#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call15.log
: > $L
echo "== full gpu suite" >> $L
timeout 900 python -m pytest tests -m gpu -q >> $L 2>&1; echo "rc=$?" >> $L
echo "== bench hy15 vae full" >> $L
timeout 300 python scripts/bench_hy15_vae.py > gpurun_out/bench_hy15_vae.json 2>> $L; echo "rc=$?" >> $L
cat gpurun_out/bench_hy15_vae.json >> $L
tail -c 3000 $L
