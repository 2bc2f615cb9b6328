#!/bin/bash
mkdir -p gpurun_out
HK_TC_DEBUG=1 HK_LIB=$PWD/heat_b200/variants/libhk_timing.so timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e > gpurun_out/t_new.log 2>&1
(cd _r1 && HK_TC_DEBUG=1 timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e > ../gpurun_out/t_r1.log 2>&1)
echo NEW; grep -A 30 "hk tc timeline" gpurun_out/t_new.log | tail -32
echo R1; grep -A 30 "hk tc timeline" gpurun_out/t_r1.log | tail -32
