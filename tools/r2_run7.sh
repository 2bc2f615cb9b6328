#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multirank.py 2>&1 | tail -15 > gpurun_out/r2_pytest7.log; tail -5 gpurun_out/r2_pytest7.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload config5 > gpurun_out/r2_c5.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload config5 --path row128 > gpurun_out/r2_c5_row128.log 2>&1
for f in r2_c5 r2_c5_row128; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/$f.log") if x.startswith("{")][-1]; j=json.loads(l)
    print("$f", round(j["ms_per_step"],3), round(j["roofline"]["kernel_ms_avg"],3), round(j["roofline"]["frac"],3), j["roofline"]["kernel"], j.get("parity"))
except Exception as e:
    print("$f", "FAILED", e); print(open("gpurun_out/$f.log").read()[-1500:])
PY
done
