#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_fuse_q_norm_bench.log
: > $L
for V in 0 1; do
  echo "== B200_WAN_FUSE_Q_NORM=$V :: bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-vae --no-reference-gpu" | tee -a $L
  B200_WAN_FUSE_Q_NORM=$V timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-vae --no-reference-gpu > gpurun_out/fq_$V.json 2> gpurun_out/fq_$V.err
  tail -2 gpurun_out/fq_$V.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/fq_$V.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','latents_sha256')}, d['roofline']['achieved'])" | tee -a $L
done
