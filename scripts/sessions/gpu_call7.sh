#!/bin/bash
# A/B of all attention variants (gated by a tiny sanity run each), then the GPU test-suite with the default.
mkdir -p gpurun_out
L=gpurun_out/call7.log
: > $L
for v in 2 4 5 6 7 8 9; do
  out=$(B200_ATTN_VARIANT=$v timeout 60 python scripts/gpu_check.py attn 1 2 300 333 2>&1 | tail -1)
  echo "-- sanity variant $v: $out" >> $L
  ok=$(python - "$out" <<'EOF'
import json, sys
try:
    r = json.loads(sys.argv[1]); print(1 if (r.get("rel_l2", 1) < 1e-2 and not r.get("nan", True)) else 0)
except Exception:
    print(0)
EOF
)
  [ "$ok" = "1" ] || continue
  for args in "1 32 1024 1024 0 42" "2 3 1000 777" "1 2 75600 75600 1" "1 40 75600 75600 1" "1 40 75600 512 1"; do
    echo "-- variant $v attn $args" >> $L
    B200_ATTN_VARIANT=$v timeout 90 python scripts/gpu_check.py attn $args >> $L 2>&1
  done
done
echo "== pytest gpu (default variant)" >> $L
timeout 600 python -m pytest tests -m gpu -q >> $L 2>&1
tail -c 1500 $L
