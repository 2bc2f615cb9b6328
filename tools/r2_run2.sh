#!/bin/bash
# round-2 GPU run 2: quick A/B of the config-3 pass on blobs / randn / uncentred (+ config 5 with the DMMA kernel)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "single_step or nan or cold_path or fit_matches" 2>&1 | tail -5
for data in blobs randn uncentred; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --data $data > gpurun_out/r2b_bench_c3_$data.log 2>&1
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload config5 > gpurun_out/r2b_bench_c5.log 2>&1
for f in c3_blobs c3_randn c3_uncentred c5; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2b_bench_$f.log") if x.startswith("{")][-1]; j=json.loads(l)
    print("$f", round(j["ms_per_step"],3), round(j["roofline"]["kernel_ms_avg"],3), round(j["roofline"]["frac"],3), j["roofline"]["kernel"], (j.get("filter") or {}).get("undecided_frac"), j.get("graph_replay"))
except Exception as e:
    print("$f", "FAILED", e); print(open("gpurun_out/r2b_bench_$f.log").read()[-1500:])
PY
done
