#!/bin/bash
# eALS re-check after a kernel change: parity tests, launch list, C3 bench line
timeout 400 python -m pytest tests/test_als_gpu.py tests/test_xl_als_gpu.py tests/test_fit_gpu.py -q 2>&1 | tail -3
bash tools/ncu_lists.sh c3 2>&1 | grep "als_\|gram" | head -12
python bench.py --workload c3 --no-cpu --no-e2e > gpurun_out/c3_new.json; python tools/show_bench.py gpurun_out/c3_new.json
