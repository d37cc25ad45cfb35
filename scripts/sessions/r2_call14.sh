#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/vae_one_tile.py 3 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 197 -c 400 --csv --log-file gpurun_out/r2_vae_tile_launches.csv python scripts/vae_one_tile.py 1 > /dev/null 2>&1
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(open('gpurun_out/r2_vae_tile_launches.csv')) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]; ki, vi = H.index('Kernel Name'), H.index('Metric Value')
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[hdr + 1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    name = re.sub(r'\(.*', '', r[ki]); tot[name] += v; cnt[name] += 1
s = sum(tot.values())
print(f"total {s/1e6:.2f} ms over {sum(cnt.values())} launches")
for k, v in tot.most_common(12): print(f"{v/1e6:8.2f} ms {100*v/s:5.1f}% x{cnt[k]:4d}  {k[:90]}")
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:linear_kernel -s 2 -c 1 -o gpurun_out/r2_linear_pair_qkv python scripts/gemm_one.py 75600 15360 5120 0 2>&1 | tail -2
