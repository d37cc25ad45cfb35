#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 1 -c 1 -f -o gpurun_out/r2_attn_final python scripts/attn_one.py 40 75600 2>&1 | tail -2
# launch list of a reduced-depth step (2 layers: the per-layer pattern repeats; shares of the kernels inside the timed region)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step.csv python bench.py --steps 1 --warmup 1 --layers 4 --no-cpu-baseline --no-vae --no-reference-gpu > /dev/null 2>&1
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(open('gpurun_out/r2_launches_step.csv')) if len(r) > 5]
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
open('gpurun_out/r2_launches_step_summary.txt', 'w').write("\n".join(out) + "\n"); print("\n".join(out))
PY
