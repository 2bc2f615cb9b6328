#!/bin/bash
# quick A/B: config-3 pass on the given datasets
mkdir -p gpurun_out
for data in "$@"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --data $data > gpurun_out/q_$data.log 2>&1
  python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/q_$data.log") if x.startswith("{")][-1]; j=json.loads(l)
    print("$data", round(j["ms_per_step"],3), round(j["roofline"]["kernel_ms_avg"],3), round(j["roofline"]["frac"],3), j["roofline"]["kernel"], (j.get("filter") or {}).get("undecided_frac"), (j.get("graph_replay") or {}).get("ms_per_step"))
except Exception as e:
    print("$data", "FAILED", e); print(open("gpurun_out/q_$data.log").read()[-1500:])
PY
done
