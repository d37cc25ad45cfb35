#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_call15_pytest.log
cat gpurun_out/r2_call15_pytest.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err || tail -5 gpurun_out/r2_bench_n1_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1_b.json'))
for k in ('value','ms_per_step','e2e','latents_sha256','gpu_launches','clocks'):
    print(k, d.get(k))
print(d['roofline']['achieved'], d['roofline']['kernel_share_of_step'], d['roofline_step']['achieved'], d['vae_decode'])
print(d.get('reference_gpu',{}).get('this_repo_over_reference_gpu'), d.get('reference_gpu',{}).get('ms_per_block'))
PY
