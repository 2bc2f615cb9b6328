#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cdist_tc_kernel --launch-skip 3 -c 1 -o gpurun_out/r02_c2_cdist -f python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-extras --workload config2 > gpurun_out/r02_c2_ncu.log 2>&1
tail -2 gpurun_out/r02_c2_ncu.log
