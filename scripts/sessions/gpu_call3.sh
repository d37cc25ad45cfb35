#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call3.log
echo "== attn A/B" > $L
for v in 1 2; do
  for args in "1 2 300 333" "1 32 1024 1024 0 42" "1 2 75600 75600 1" "1 40 75600 75600 1" "1 40 75600 512 1"; do
    echo "-- variant $v attn $args" >> $L
    B200_ATTN_VARIANT=$v timeout 300 python scripts/gpu_check.py attn $args >> $L 2>&1
  done
done
echo "== pytest gpu" >> $L
timeout 1200 python -m pytest tests -m gpu -q >> $L 2>&1
echo "== ncu full attention v2" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 0 -c 1 -o gpurun_out/prof_attn \
   python scripts/gpu_check.py attn 1 40 75600 75600 0 >> $L 2>&1
echo "== ncu launch list (timed region of bench, 1 step)" >> $L
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2>> $L
echo "== bench full" >> $L
timeout 1200 python bench.py > gpurun_out/bench_full.json 2>> $L
cat gpurun_out/bench_full.json >> $L
tail -c 3000 $L
