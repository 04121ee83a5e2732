#!/bin/bash
# the driver's round-end scaling launch at N = 2 (both arms), plus the two tests that failed in the last full run
timeout 300 python -m pytest tests/test_ncf.py -q 2>&1 | tail -3
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}"; }
timeout 400 bash -c "$(declare -f run); run 29621 --impl reference --steps 3 --warmup 1" > gpurun_out/scale_ref_n2.json 2> gpurun_out/scale_ref_n2.err; tail -2 gpurun_out/scale_ref_n2.err; python tools/show_bench.py gpurun_out/scale_ref_n2.json
timeout 400 bash -c "$(declare -f run); run 29622 --steps 20 --warmup 5" > gpurun_out/scale_n2.json 2> gpurun_out/scale_n2.err; tail -2 gpurun_out/scale_n2.err; python tools/show_bench.py gpurun_out/scale_n2.json
