#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call13.log
: > $L
B200_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/flux_launches.csv python scripts/bench_flux.py --steps 1 --warmup 2 >> $L 2>&1; echo "rc=$?" >> $L
python scripts/launch_share.py gpurun_out/flux_launches.csv >> $L 2>&1
tail -c 3500 $L
