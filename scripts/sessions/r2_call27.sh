#!/bin/bash
# full GPU suite + smoke + headline bench with the persistent / staged attention kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_call27_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r2_call27_smoke.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_n1_persist.json 2> gpurun_out/r2_bench_n1_persist.err
tail -c 3000 gpurun_out/r2_bench_n1_persist.json
