#!/bin/bash
# ncu --set full of the lane-group row kernels of one C3 half-sweep (user side)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'als_rows_group_blocked_kernel' -c 5 -o gpurun_out/als_group_r2 -f python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu --no-e2e --no-also > /dev/null 2>&1
ls -la gpurun_out/als_group_r2.ncu-rep
