#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_lora.py tests/test_gpu_baseline_width.py tests/test_scheduler.py -m gpu -x -q -s 2>&1 | tail -15 > gpurun_out/r2_call13_pytest.log
cat gpurun_out/r2_call13_pytest.log
timeout 900 python bench.py --steps 2 --warmup 3 --cuda-graph --no-cpu-baseline --no-vae > gpurun_out/r2_bench_n1_graph.json 2> gpurun_out/r2_bench_n1_graph.err || tail -5 gpurun_out/r2_bench_n1_graph.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1_graph.json'))
for k in ('value','ms_per_step','e2e','latents_sha256','gpu_launches','reference_gpu'):
    print(k, d.get(k))
print(d['roofline']['achieved'], d['roofline_step']['achieved'], d['config'])
PY
