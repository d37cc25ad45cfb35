#!/bin/bash
# final single-GPU validation of the round-1 tree
mkdir -p gpurun_out
L=gpurun_out/call30.log
: > $L
timeout 200 python __graft_entry__.py smoke >> $L 2>&1; echo "rc=$?" >> $L
timeout 900 python -m pytest tests -m gpu -q >> $L 2>&1; echo "rc=$?" >> $L
tail -c 1500 $L
