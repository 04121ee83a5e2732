#!/bin/bash
# One multi-GPU gpurun call (charged N x box time): the W = 2/4/8 tests, then the bench lines at N ranks.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1500 -- 'bash tools/multi_gpu_run.sh 8'
N=${1:-8}
O=gpurun_out/multi
mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
echo "== dist tests (BPR W=2,4,8: replicas identical, sharded Evaluate, NDCG vs W=1; eALS W=2,4)"
timeout 600 python -m pytest tests/test_dist_gpu.py -q -s -k "${TESTS:-4 or 8}" 2>&1 | grep -v "^$" | tail -12
echo "== c5 (BASELINE configs[4]) on $N ranks"
timeout 600 bash -c "$(declare -f run); N=$N; run 29611 --workload c5 --steps 10 --warmup 3" > $O/c5_n$N.json 2> $O/c5_n$N.err; tail -2 $O/c5_n$N.err; python tools/show_bench.py $O/c5_n$N.json
if [ -n "$WITH_C2" ]; then
echo "== c2 weak scaling on $N ranks (what the driver's SCALE run does at round end)"
timeout 400 bash -c "$(declare -f run); N=$N; run 29612 --steps 20 --warmup 5" > $O/c2_n$N.json 2> $O/c2_n$N.err; tail -2 $O/c2_n$N.err; python tools/show_bench.py $O/c2_n$N.json
fi
echo "== c4 top-k, queries sharded over $N ranks"
timeout 400 bash -c "$(declare -f run); N=$N; run 29613 --workload c4 --steps 5 --warmup 3 --no-e2e" > $O/c4_n$N.json 2> $O/c4_n$N.err; tail -2 $O/c4_n$N.err; python tools/show_bench.py $O/c4_n$N.json
