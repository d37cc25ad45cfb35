#!/bin/bash
# smoke + GPU tests + bench + ncu launch list + ncu full capture of the attention kernel
mkdir -p gpurun_out
echo "== smoke" > gpurun_out/call2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/call2.log 2>&1
echo "== pytest gpu" >> gpurun_out/call2.log
timeout 900 python -m pytest tests -m gpu -x -q >> gpurun_out/call2.log 2>&1
echo "== bench layers=2 (sanity)" >> gpurun_out/call2.log
timeout 600 python bench.py --steps 1 --warmup 1 --layers 2 --no-cpu-baseline > gpurun_out/bench_l2.json 2>> gpurun_out/call2.log
cat gpurun_out/bench_l2.json >> gpurun_out/call2.log
echo "== bench full" >> gpurun_out/call2.log
timeout 1200 python bench.py > gpurun_out/bench_full.json 2>> gpurun_out/call2.log
cat gpurun_out/bench_full.json >> gpurun_out/call2.log
echo "== ncu launch list" >> gpurun_out/call2.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2>> gpurun_out/call2.log
echo "== ncu full attention" >> gpurun_out/call2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 1 -c 1 -o gpurun_out/prof_attn \
   python scripts/gpu_check.py attn 1 40 75600 75600 0 >> gpurun_out/call2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_kernel -s 1 -c 1 -o gpurun_out/prof_linear \
   python scripts/gpu_check.py linear 75600 5120 5120 0 1 >> gpurun_out/call2.log 2>&1
tail -c 6000 gpurun_out/call2.log
