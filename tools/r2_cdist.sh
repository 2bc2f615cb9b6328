#!/bin/bash
# cdist A/B: parity tests on the new library, then config 2 with the previous and the new build
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 300 python -m pytest tests -m gpu -x -q -k "cdist or bigk" > gpurun_out/cdist_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/cdist_tests.log; fi
tail -3 gpurun_out/cdist_tests.log
for v in "$@"; do
  if [ "$v" = main ]; then lib=""; else lib=$PWD/heat_b200/variants/libhk_$v.so; fi
  HK_LIB=$lib timeout 200 python bench.py --workload config2 --no-cpu --no-e2e --no-extras --steps 20 --warmup 5 > gpurun_out/cdist_$v.json 2> gpurun_out/cdist_$v.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/cdist_$v.json").read().strip().splitlines()[-1])
    print("$v", j["ms_per_step"], j["roofline"]["frac"], j.get("clocks",{}).get("sm_mhz"))
except Exception as e:
    print("$v failed", e)
PY
done
