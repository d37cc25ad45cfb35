#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n2_final.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus', 'gpu_launches', 'latents_sha256', 'frames_per_sec')}, d['config'].get('parallelism'), d['roofline']['achieved'], d['e2e']['value'], d['vae_decode']['ms'], d['vae_decode']['frames_sha256'][:12], d['clocks']['reasons'])
PY
tail -2 gpurun_out/r2_bench_n2_final.err
