#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "cold_path or nan" 2>&1 | tail -5
for d in blobs randn uncentred; do bash tools/r2_variants.sh $d default nohelp; done
