#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:lloyd_tc_kernel --launch-skip 2 -c 1 -o gpurun_out/r2d_tc_new -f python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2d_ncu.log 2>&1
(cd _r1 && timeout 900 ncu --set full --import-source on --clock-control none -k regex:lloyd_tc_kernel --launch-skip 2 -c 1 -o ../gpurun_out/r2d_tc_r1 -f python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > ../gpurun_out/r2d_ncu_r1.log 2>&1)
ls -la gpurun_out/*.ncu-rep | tail -3
