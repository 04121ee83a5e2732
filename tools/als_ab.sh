O=gpurun_out/round_check; mkdir -p $O
echo "== ALS parity default"; timeout 60 python -m pytest tests/test_als_gpu.py -x -q 2>&1 | tail -2
echo "== ALS parity g16"; GORSE_B200_ALS_G16=1 timeout 60 python -m pytest tests/test_als_gpu.py -x -q 2>&1 | tail -2
timeout 60 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu > $O/bench_c3_b1_wide.json 2> $O/bench_c3.err
GORSE_B200_ALS_G16=1 timeout 60 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu > $O/bench_c3_b1_g16.json 2>> $O/bench_c3.err
timeout 60 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu > $O/bench_c3_b1_wide_again.json 2>> $O/bench_c3.err
for f in $O/bench_c3_b1_wide.json $O/bench_c3_b1_g16.json $O/bench_c3_b1_wide_again.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['ms_per_step'],2))"; done
