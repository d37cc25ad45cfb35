#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call21.log
: > $L
timeout 200 python scripts/row_kernels_bw.py > gpurun_out/row_kernels_bw.json 2>> $L; echo "rc=$?" >> $L
B200_PROFILE=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
   --clock-control none --csv --log-file gpurun_out/row_kernels_ncu.csv python scripts/row_kernels_bw.py >> $L 2>&1; echo "rc=$?" >> $L
tail -c 2500 $L
