#!/bin/bash
for st in 4 5; do
  echo "== GORSE_B200_TOPK_STAGES=$st"
  GORSE_B200_TOPK_STAGES=$st timeout 300 python bench.py --workload c4 --no-cpu --no-e2e --steps 5 > gpurun_out/c4_st$st.json 2>gpurun_out/c4_st$st.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/c4_st$st.json').read().strip().splitlines()[-1])
print("  ms/step %.1f  value %.3g  stage1 ms %.1f frac %.3f  fallback rows %d" % (d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['fallback_rows']), d['clocks'])
PY
done
