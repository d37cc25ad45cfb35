#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_call16.log
: > $L
run() { echo "== $1" >> $L; env $1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-vae --no-reference-gpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({'ms_per_step': round(d['ms_per_step'],1), 'attn_tflops': round(d['roofline']['achieved'],1), 'attn_share': round(d['roofline']['kernel_share_of_step'],4), 'non_attn_ms': round(d['ms_per_step']*(1-d['roofline']['kernel_share_of_step']),1), 'sm_mhz': d['clocks']['sm_mhz'], 'sha': d['latents_sha256'][:12]})" >> $L 2>&1; }
run "B200_LINEAR_QUAD=1"
run "B200_LINEAR_QUAD=0"
run "B200_LINEAR_QUAD=0 B200_LINEAR_GROUP_M=16 B200_LINEAR_PANEL_N=9999"
run "B200_LINEAR_QUAD=1 B200_LINEAR_GROUP_M=16 B200_LINEAR_PANEL_N=9999"
run "B200_LINEAR_QUAD=1"
cat $L
