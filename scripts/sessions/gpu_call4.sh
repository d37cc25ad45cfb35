#!/bin/bash
# Gated GPU session: cheap sanity first, every step under its own short timeout, whole script bounded by the caller.
mkdir -p gpurun_out
L=gpurun_out/call4.log
: > $L
ok_attn() {  # $1 = variant ; prints 1 if a tiny attention call returns a sane result within 90 s
  out=$(B200_ATTN_VARIANT=$1 timeout 90 python scripts/gpu_check.py attn 1 2 300 333 2>&1 | tail -1)
  echo "-- sanity variant $1: $out" >> $L
  python - "$out" <<'EOF'
import json, sys
try:
    r = json.loads(sys.argv[1]); print(1 if (r.get("rel_l2", 1) < 1e-2 and not r.get("nan", True)) else 0)
except Exception:
    print(0)
EOF
}
OK1=$(ok_attn 1); OK2=$(ok_attn 2); OK3=$(ok_attn 3)
V=0
[ "$OK1" = "1" ] && V=1
[ "$OK2" = "1" ] && V=2
echo "== sanity: v1=$OK1 v2=$OK2 v3=$OK3 -> using $V" >> $L
export B200_ATTN_VARIANT=$V

if [ "$V" != "0" ]; then
  echo "== attn A/B" >> $L
  for v in 1 2 3; do
    eval okv=\$OK$v
    [ "$okv" = "1" ] || continue
    for args in "1 32 1024 1024 0 42" "1 2 75600 75600 1" "1 40 75600 75600 1" "1 40 75600 512 1"; do
      echo "-- variant $v attn $args" >> $L
      B200_ATTN_VARIANT=$v timeout 120 python scripts/gpu_check.py attn $args >> $L 2>&1
    done
  done
  echo "== pytest gpu (all)" >> $L
  timeout 900 python -m pytest tests -m gpu -q >> $L 2>&1
else
  echo "== pytest gpu (no attention)" >> $L
  timeout 900 python -m pytest tests -m gpu -q -k "not attention and not dit and not denoise" >> $L 2>&1
fi

if [ "$V" != "0" ]; then
  echo "== ncu full attention" >> $L
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 0 -c 1 -o gpurun_out/prof_attn \
     python scripts/gpu_check.py attn 1 40 75600 75600 0 >> $L 2>&1
  echo "== bench full" >> $L
  timeout 700 python bench.py > gpurun_out/bench_full.json 2>> $L
  cat gpurun_out/bench_full.json >> $L
  echo "== ncu launch list (timed region of bench, 1 step)" >> $L
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
     --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-vae \
     > gpurun_out/bench_under_ncu.json 2>> $L
fi
tail -c 2500 $L
