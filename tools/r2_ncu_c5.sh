#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:lloyd_dmma_kernel --launch-skip 2 -c 1 -o gpurun_out/r2_c5_dmma -f python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --workload config5 > gpurun_out/r2_c5_ncu.log 2>&1
ls -la gpurun_out/r2_c5_dmma.ncu-rep
