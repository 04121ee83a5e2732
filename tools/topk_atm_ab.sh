#!/bin/bash
for mode in 1 0; do
  export GORSE_B200_TOPK_ATM=$mode
  echo "== GORSE_B200_TOPK_ATM=$mode"
  timeout 300 python -m pytest tests/test_topk_mma_gpu.py tests/test_topk_gpu.py tests/test_logics_gpu.py -q -x 2>&1 | tail -4
  timeout 300 python bench.py --workload c4 --no-cpu --no-e2e --steps 5 > gpurun_out/c4_atm$mode.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('gpurun_out/c4_atm$mode.json').read().strip().splitlines()[-1])
print("  ms/step %.1f  value %.3g  stage1 ms %.1f frac %.3f  fallback rows %d" % (d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['fallback_rows']))
PY
done
