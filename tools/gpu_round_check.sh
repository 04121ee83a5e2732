#!/bin/bash
# One gpurun call that re-validates the whole GPU side after a change (kept short: box time is budgeted).
#   /usr/local/graft/bin/gpurun --timeout 420 -- 'bash tools/gpu_round_check.sh'
# Everything lands in gpurun_out/round_check/ ; the last lines of each step are echoed.
O=gpurun_out/round_check
mkdir -p $O
step() { echo "== $1"; }
step "ALS parity, four coordinates per butterfly"
GORSE_B200_ALS_BLOCK=4 timeout 60 python -m pytest tests/test_als_gpu.py -x -q 2>&1 | tail -2
step "ALS parity, whole-warp classes"
GORSE_B200_ALS_G16=0 timeout 60 python -m pytest tests/test_als_gpu.py -x -q 2>&1 | tail -2
step "ALS C3 bench: default / block=4 / whole-warp classes"
timeout 60 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu > $O/bench_c3.json 2> $O/bench_c3.err
GORSE_B200_ALS_BLOCK=4 timeout 60 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu > $O/bench_c3_block4.json 2>> $O/bench_c3.err
GORSE_B200_ALS_G16=0 timeout 60 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu > $O/bench_c3_wide.json 2>> $O/bench_c3.err
for f in $O/bench_c3.json $O/bench_c3_block4.json $O/bench_c3_wide.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/epoch", round(d["ms_per_step"], 2), "frac", round(d["roofline"]["frac"], 4))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
step "full GPU suite"
timeout 240 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
step "smoke"
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
step "bench c2 (default)"
timeout 150 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 600 $O/bench_c2.json
step "bench c4"
timeout 120 python bench.py --workload c4 > $O/bench_c4.json 2> $O/bench_c4.err; tail -c 400 $O/bench_c4.json
step "reference arm"
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 300 $O/bench_ref.json
