#!/bin/bash
# multi-GPU session (run with gpurun --gpus N): correctness of the sharded paths, then the bench at N
N=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/call5_n$N.log
: > $L
echo "== mp check N=$N" >> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   scripts/gpu_mp_check.py >> $L 2>&1
echo "== bench N=$N (NCCL exchange)" >> $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_n$N.json 2>> $L
cat gpurun_out/bench_n$N.json >> $L
if [ "$N" != "2" ]; then
  echo "== bench N=$N (fused peer-memory exchange)" >> $L
  B200_P2P=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
     bench.py --gpus $N --steps 2 --warmup 3 --no-vae > gpurun_out/bench_n${N}_p2p.json 2>> $L
  cat gpurun_out/bench_n${N}_p2p.json >> $L
fi
tail -c 3000 $L
