#!/bin/bash
# per-launch durations (ncu --metrics gpu__time_duration.sum --clock-control none) of one short bench run per workload
O=gpurun_out/ncu
mkdir -p $O
for wl in "$@"; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$wl.csv \
      python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --no-e2e --no-also > $O/bench_$wl.log 2>&1
  python tools/summarize_ncu.py list $O/launches_$wl.csv $O/launches_$wl.md "$wl launch list"; head -50 $O/launches_$wl.md
done
