#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_attn_cross_timeline2.log
: > $L
echo "== timeline persistent staged 40 x 75600 x 512" | tee -a $L
timeout 120 python scripts/attn_timeline_items.py 40 75600 512 20 2 2>&1 | tee -a $L
