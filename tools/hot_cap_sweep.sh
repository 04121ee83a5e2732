#!/bin/bash
# BPR hot-row concurrency cap: throughput + 100-epoch stability (finite factors, NDCG of the fit) per cap, Zipf 1.0 and 1.3
O=gpurun_out/hotcap
mkdir -p $O
for z in 1.0 1.3; do
  for cap in 768 1536 3072; do
    GORSE_B200_HOT_ROW_CONCURRENCY=$cap timeout 200 python bench.py --zipf $z --steps 10 --warmup 3 --no-cpu --no-also > $O/c2_z${z}_cap$cap.json 2> $O/err.log
    echo "zipf $z cap $cap"; python tools/show_bench.py $O/c2_z${z}_cap$cap.json
  done
done
