#!/bin/bash
# timing experiments on the sweep kernel (results are garbage with EXP != 0: only the first launch is timed, then the run is killed)
for x in 0 4 1 2; do
  echo "== GORSE_B200_TOPK_EXP=$x"
  GORSE_B200_TOPK_EXP=$x timeout 300 ncu --csv --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max -k regex:topk_mma_kernel -c 1 --kill yes --clock-control none python bench.py --workload c4 --steps 1 --warmup 0 --no-cpu --no-e2e --no-also > gpurun_out/exp_$x.csv 2>&1
  grep "topk_mma_kernel" gpurun_out/exp_$x.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
