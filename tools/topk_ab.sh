#!/bin/bash
for cfg in "3.2 4" "3.2 1" "4.5 4" "4.5 1" "6 1"; do
  set -- $cfg
  GORSE_B200_TOPK_MARGIN=$1 GORSE_B200_TOPK_CHUNK=$2 python bench.py --workload c4 --no-cpu --no-e2e --steps 5 > gpurun_out/c4_ab.json 2>/dev/null
  echo "margin $1 chunk x$2"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c4_ab.json').read().strip().splitlines()[-1])
print("  ms/step %.1f  value %.3g  stage1 ms %.1f frac %.3f  fallback rows (all calls) %d" % (d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['fallback_rows']))
PY
done
