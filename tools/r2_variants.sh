#!/bin/bash
# A/B of library variants built with tools/build_variant.sh: tools/r2_variants.sh DATA name1 name2 ...
data=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  lib=""; [ "$v" != "default" ] && lib="$PWD/heat_b200/variants/libhk_$v.so"
  HK_LIB=$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --data $data > gpurun_out/v_${v}_$data.log 2>&1
  python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/v_${v}_$data.log") if x.startswith("{")][-1]; j=json.loads(l)
    print("$v $data", round(j["ms_per_step"],3), round(j["roofline"]["kernel_ms_avg"],3), round(j["roofline"]["frac"],3), j["roofline"]["kernel"], (j.get("filter") or {}).get("undecided_frac"))
except Exception as e:
    print("$v $data", "FAILED", e); print(open("gpurun_out/v_${v}_$data.log").read()[-1500:])
PY
done
