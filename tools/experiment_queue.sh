#!/bin/bash
# Experiments that were written without GPU time left (end of round 1), all OFF by default.  One gpurun call runs the
# parity tests and the bench under each switch so that the next round can flip (or delete) them on evidence:
#   /usr/local/graft/bin/gpurun --timeout 500 -- 'bash tools/experiment_queue.sh'
#   GORSE_B200_ALS_FMA2=1   Gram register tiles issue FFMA2 (2 fp32 FMAs / instruction): gram_kernel, als_chunk_gram_kernel
#   GORSE_B200_TOPK_EPI8=1  stage-1 epilogue tests 8 columns per branch with a 3-input max tree (FMNMX3)
#   GORSE_B200_HOT_PREFETCH=1  BPR capped hot apply prefetches the user row two rounds ahead into L2 (round time is set by the
#                           HBM latency of that gather: ~900 rounds x 1.9 us for the top item at C2)
#   GORSE_B200_ALS_BLOCK=4  four coordinates per shuffle butterfly (measured slower in round 1; kept for reference)
#   GORSE_B200_ALS_G16=0    whole-warp classes for 9..32-entry rows (measured slower in round 1)
O=gpurun_out/experiments
mkdir -p $O
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 3), "value", f'{d["value"]:.4g}', "frac", round(d["roofline"]["frac"], 4))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
echo "== HOT_PREFETCH: BPR parity (atomic paths) + C2 bench (baseline first)"
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > $O/c2_base.json 2> $O/c2.err; line $O/c2_base.json
GORSE_B200_HOT_PREFETCH=1 timeout 120 python -m pytest tests/test_bpr_gpu.py tests/test_fit_gpu.py -x -q 2>&1 | tail -2
GORSE_B200_HOT_PREFETCH=1 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > $O/c2_prefetch.json 2>> $O/c2.err; line $O/c2_prefetch.json
GORSE_B200_HOT_PREFETCH=1 timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu > $O/c2_prefetch_e2e.json 2>> $O/c2.err   # 100-epoch stability leg
python -c "import json;d=json.loads(open('$O/c2_prefetch_e2e.json').read().strip().splitlines()[-1]);print('e2e',d.get('e2e'),d.get('e2e_error'))"
echo "== FMA2: ALS parity + C3 bench (baseline first)"
timeout 90 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu > $O/c3_base.json 2> $O/c3.err; line $O/c3_base.json
GORSE_B200_ALS_FMA2=1 timeout 90 python -m pytest tests/test_als_gpu.py -x -q 2>&1 | tail -2
GORSE_B200_ALS_FMA2=1 timeout 90 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu > $O/c3_fma2.json 2>> $O/c3.err; line $O/c3_fma2.json
echo "== EPI8: top-k parity + C4 bench (baseline first)"
timeout 150 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu > $O/c4_base.json 2> $O/c4.err; line $O/c4_base.json
GORSE_B200_TOPK_EPI8=1 timeout 150 python -m pytest tests/test_topk_mma_gpu.py tests/test_topk_gpu.py tests/test_logics_gpu.py -x -q 2>&1 | tail -2
GORSE_B200_TOPK_EPI8=1 timeout 150 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu > $O/c4_epi8.json 2>> $O/c4.err; line $O/c4_epi8.json
echo "== experimental sparse Dot index (csrc/sparse.cu): parity vs the oracle + reference known answers"
GORSE_B200_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_sparse_gpu.py -x -q 2>&1 | tail -3
# separately, on two GPUs:  gpurun --gpus 2 -- 'python -m pytest tests/test_dist_gpu.py -x -q'   (BPR exchange + the multi-rank eALS epoch)
