#!/bin/bash
# One gpurun call that re-validates the whole GPU side after a change (kept short: box time is budgeted).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_round_check.sh'
# Everything lands in gpurun_out/round_check/ ; the last lines of each step are echoed.
O=gpurun_out/round_check
mkdir -p $O
echo "== full GPU suite"
timeout 500 python -m pytest tests -q -m gpu 2>&1 | tail -15
echo "== smoke"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (default: c2 + also c3, c4)"
timeout 400 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.err; python tools/show_bench.py $O/bench_default.json
echo "== reference arm"
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; python tools/show_bench.py $O/bench_ref.json
if [ -n "$WITH_LISTS" ]; then
echo "== ncu launch lists (c2 c3 c4)"
bash tools/ncu_lists.sh c2 c3 c4 > $O/lists.log 2>&1; tail -3 $O/lists.log
fi
