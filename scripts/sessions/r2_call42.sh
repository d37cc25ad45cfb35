#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29599 bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_bench_n8_s2.json 2> gpurun_out/r2_bench_n8_s2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n8_s2.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus', 'gpu_launches', 'latents_sha256', 'frames_per_sec')}, d['config'].get('parallelism'), d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['value'], d['vae_decode'], d['clocks'])
PY
tail -3 gpurun_out/r2_bench_n8_s2.err
