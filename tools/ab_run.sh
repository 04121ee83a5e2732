#!/bin/bash
# one GPU call: top-k epilogue check (early release + merged threshold) and margin sweep
echo "== topk"
timeout 400 python -m pytest tests/test_topk_mma_gpu.py tests/test_topk_gpu.py tests/test_logics_gpu.py tests/test_fullsize_gpu.py tests/test_vecdb_gpu.py -q 2>&1 | tail -4
for mg in 3.2 4; do
  export GORSE_B200_TOPK_MARGIN=$mg
  echo "== margin $mg"
  timeout 300 python bench.py --workload c4 --no-cpu --no-e2e --steps 5 > gpurun_out/c4_m$mg.json 2>gpurun_out/c4_m$mg.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/c4_m$mg.json').read().strip().splitlines()[-1])
print("  ms/step %.1f  value %.3g  stage1 ms %.1f frac %.3f  fallback rows %d" % (d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['fallback_rows']))
PY
done
unset GORSE_B200_TOPK_MARGIN
bash tools/ncu_lists.sh c4 2>&1 | grep "topk\|prune\|exact" | head
timeout 600 ncu --set full --clock-control none --import-source on -k regex:topk_mma_kernel -s 1 -c 1 -o gpurun_out/topk_r2_tree -f python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu --no-e2e --no-also > /dev/null 2>&1
