#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_attn_cross_timeline.log
: > $L
echo "== persistent, 40 x 75600 x 512" | tee -a $L
timeout 120 python scripts/attn_timeline_items.py 40 75600 512 20 3 2>&1 | tee -a $L
echo "== one CTA per item, 40 x 75600 x 512 (first item only)" | tee -a $L
B200_ATTN_PERSIST=0 timeout 120 python scripts/attn_timeline_items.py 40 75600 512 0 1 2>&1 | tee -a $L
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 1 -f -o gpurun_out/r2_attn_cross_persist python scripts/attn_one.py 40 75600 512 2>&1 | tail -2
