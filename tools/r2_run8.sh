#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 100 python tools/dbg_c5.py nothing 2>&1 | grep -E "HANG|ALL OK 2"; done
timeout -s ABRT 200 python -X faulthandler bench.py --workload config5 --steps 10 --warmup 3 --ref-gpu-rows 0 > gpurun_out/r02_bench_config5.log 2>&1; echo "c5 full rc=$?"; grep "^{" gpurun_out/r02_bench_config5.log | tail -1 > gpurun_out/r02_bench_config5.json
python - <<PY
import json
j=json.load(open("gpurun_out/r02_bench_config5.json")); print("c5", round(j["value"],1), j["ms_per_step"], j["roofline"]["kernel_ms_avg"], round(j["roofline"]["frac"],3), j["roofline"]["kernel"], (j.get("e2e") or {}).get("value"), (j.get("cpu_baseline") or {}).get("value"))
PY
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multirank.py 2>&1 | tail -4
bash tools/r2_variants.sh blobs default
