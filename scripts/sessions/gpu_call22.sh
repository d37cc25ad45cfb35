#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call22.log
: > $L
timeout 600 python -m pytest tests/test_gpu_hy15_vae.py tests/test_gpu_flux.py tests/test_gpu_hy15.py tests/test_gpu_qwen.py tests/test_gpu_flux2.py -m gpu -q >> $L 2>&1; echo "rc=$?" >> $L
timeout 200 python scripts/row_kernels_bw.py > gpurun_out/row_kernels_bw.json 2>> $L; echo "rc=$?" >> $L
timeout 200 python scripts/bench_hy15_vae.py --tiles 16 >> $L 2>&1; echo "rc=$?" >> $L
tail -c 3000 $L
