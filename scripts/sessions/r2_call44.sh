#!/bin/bash
# final tree: full GPU suite + smoke + headline bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_call44_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/r2_call44_smoke.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_n1_s2final2.json 2> gpurun_out/r2_bench_n1_s2final2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n1_s2final2.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'latents_sha256', 'frames_per_sec')}, d['roofline']['achieved'], d['roofline']['frac'], d['roofline_step']['frac_per_gpu'], d['e2e']['value'], d['vae_decode'], d['clocks'], d['reference_gpu']['this_repo_over_reference_gpu'], d['cpu_baseline']['value'])
PY
