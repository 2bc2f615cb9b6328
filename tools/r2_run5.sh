#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "cold_path or nan" 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "single_step or fit_matches or full_size" 2>&1 | tail -8
bash tools/r2_variants.sh blobs default; bash tools/r2_variants.sh randn default; bash tools/r2_variants.sh uncentred default
