#!/bin/bash
# round-2 GPU run 1: full GPU test-suite + config-3 bench on the four input distributions
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multirank.py 2>&1 | tail -40 > gpurun_out/r2_pytest1.log
for data in blobs randn uncentred overlap; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --data $data > gpurun_out/r2_bench_c3_$data.log 2>&1
done
tail -3 gpurun_out/r2_pytest1.log
for data in blobs randn uncentred overlap; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2_bench_c3_$data.log") if x.startswith("{")][-1]; j=json.loads(l)
    print("$data", round(j["ms_per_step"],3), round(j["roofline"]["kernel_ms_avg"],3), round(j["roofline"]["frac"],3), j.get("filter"), j.get("graph_replay"), j.get("parity"))
except Exception as e:
    print("$data", "FAILED", e); print(open("gpurun_out/r2_bench_c3_$data.log").read()[-1500:])
PY
done
