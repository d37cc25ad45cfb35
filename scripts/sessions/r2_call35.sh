#!/bin/bash
# ncu full captures of the new attention kernels (self: single-item kernel; cross: persistent staged) + variant tests +
# the image / video model step benches with clocks
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 1 -c 1 -f -o gpurun_out/r2_attn_self_final python scripts/attn_one.py 40 75600 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 1 -f -o gpurun_out/r2_attn_cross_final python scripts/attn_one.py 40 75600 512 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_variants.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_call35_variants.log
timeout 600 python scripts/bench_flux.py --graph > gpurun_out/r2_bench_flux_dev_n1.json 2> gpurun_out/r2_bench_flux.err; tail -c 1500 gpurun_out/r2_bench_flux_dev_n1.json
timeout 600 python scripts/bench_qwen.py > gpurun_out/r2_bench_qwen_edit_n1.json 2> gpurun_out/r2_bench_qwen.err; tail -c 900 gpurun_out/r2_bench_qwen_edit_n1.json
timeout 900 python scripts/bench_hy15.py > gpurun_out/r2_bench_hy15_n1.json 2> gpurun_out/r2_bench_hy15.err; tail -c 900 gpurun_out/r2_bench_hy15_n1.json
