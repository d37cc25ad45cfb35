#!/bin/bash
# validation of the tree: full GPU suite + smoke + headline bench + launch list of a reduced-depth step
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_call41_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/r2_call41_smoke.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_n1_s2final.json 2> gpurun_out/r2_bench_n1_s2final.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n1_s2final.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'latents_sha256', 'frames_per_sec')}, d['roofline']['achieved'], d['roofline']['frac'], d['roofline_step']['frac_per_gpu'], d['e2e']['value'], d['vae_decode'], d['clocks'], d['reference_gpu']['this_repo_over_reference_gpu'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step_s2.csv python bench.py --steps 1 --warmup 1 --layers 4 --no-cpu-baseline --no-vae --no-reference-gpu > /dev/null 2>&1
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(open('gpurun_out/r2_launches_step_s2.csv')) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]; ki, vi = H.index('Kernel Name'), H.index('Metric Value')
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[hdr + 1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    name = re.sub(r'\(.*', '', r[ki]); name = re.sub(r'^void ', '', name); tot[name] += v; cnt[name] += 1
s = sum(tot.values())
out = [f"ncu launch list of the timed region of `bench.py --steps 1 --layers 4` (cold-cache, serialised launches: compare SHARES): total {s/1e6:.1f} ms over {sum(cnt.values())} launches"]
for k, v in tot.most_common(14): out.append(f"{v/1e6:9.2f} ms {100*v/s:5.1f}% x{cnt[k]:4d}  {k[:100]}")
open('gpurun_out/r2_launches_step_s2_summary.txt', 'w').write("\n".join(out) + "\n"); print("\n".join(out))
PY
