#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_gemm_epi_ab.log
: > $L
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "linear" 2>&1 | tail -2 | tee -a $L
for V in "APEX_B200_LIB=$PWD/apex-studio_b200/libapex_b200_oldlinear.so" "X=1" "APEX_B200_LIB=$PWD/apex-studio_b200/libapex_b200_oldlinear.so" "X=1"; do
  env $V timeout 300 python scripts/gemm_epi_ab.py 2>&1 | tail -1 | tee -a $L
done
