#!/bin/bash
mkdir -p gpurun_out
for d in "$@"; do
HK_TC_DEBUG=1 HK_LIB=$PWD/heat_b200/variants/libhk_${VAR:-timing}.so timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --data $d > gpurun_out/t_$d.log 2>&1
echo "== $d"; grep -B1 -A10 "tiles/CTA 5279" gpurun_out/t_$d.log | tail -12
done
