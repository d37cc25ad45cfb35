#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_call12_pytest.log
cat gpurun_out/r2_call12_pytest.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1_a.json'))
for k in ('value','ms_per_step','roofline','roofline_step','e2e','clocks','latents_sha256','vae_decode','gpu_launches'):
    print(k, d.get(k))
PY
