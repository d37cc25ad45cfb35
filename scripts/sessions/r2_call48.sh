#!/bin/bash
# norm_q of the cross-attention folded into projection + logits (12 launches per block): unit test, Wan parity, bench A/B
mkdir -p gpurun_out
L=gpurun_out/r2_fuse_q_norm.log
: > $L
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_width.py -x -q -m gpu 2>&1 | tail -4 | tee -a $L
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1 | tee -a $L
for V in 0 1; do
  echo "== B200_WAN_FUSE_Q_NORM=$V :: bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-vae --no-reference-gpu" | tee -a $L
  B200_WAN_FUSE_Q_NORM=$V timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-vae --no-reference-gpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','latents_sha256')}, d['roofline']['achieved'])" | tee -a $L
done
