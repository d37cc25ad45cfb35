#!/bin/bash
mkdir -p gpurun_out
timeout 500 python scripts/gemm_group_m.py > gpurun_out/call28.log 2>&1
cat gpurun_out/call28.log
