#!/bin/bash
mkdir -p gpurun_out
for TB in 0 4; do
  B200_CONV_TB=$TB timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none --csv --log-file gpurun_out/r2_vae_tile_tb$TB.csv python scripts/vae_one_tile.py 1 > /dev/null 2>&1
done
python - <<'PY'
import csv, collections, re
for TB in (0, 4):
    rows = [r for r in csv.reader(open(f'gpurun_out/r2_vae_tile_tb{TB}.csv')) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    H = rows[hdr]; ki, vi, mi, gi, idi = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Name'), H.index('Grid Size'), H.index('ID')
    per = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        try: v = float(r[vi].replace(',', ''))
        except ValueError: continue
        per.setdefault(r[idi], {'name': re.sub(r'\(.*', '', r[ki]).replace('void ', ''), 'grid': r[gi]})[r[mi]] = v
    ids = list(per)
    half = ids[len(ids)//2:]   # second decode (the first is the warm-up)
    agg = collections.OrderedDict()
    for i in half:
        d = per[i]; k = (d['name'][:40], d['grid'])
        a = agg.setdefault(k, [0, 0.0, 0.0, 0.0, 0.0]); a[0] += 1; a[1] += d.get('gpu__time_duration.sum', 0)
        a[2] += d.get('sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_active', 0)
        a[3] += d.get('lts__t_bytes.sum', 0); a[4] += d.get('l1tex__m_xbar2l1tex_read_bytes.sum', 0)
    tot = sum(a[1] for a in agg.values())
    print(f"== TB={TB}: total {tot/1e6:.2f} ms")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        print(f"{a[1]/1e6:8.3f} ms x{a[0]:3d} tensor {a[2]/a[0]:5.1f}%  lts {a[3]/1e9:7.2f} GB  xbar->l1 {a[4]/1e9:7.2f} GB  {k}")
PY
