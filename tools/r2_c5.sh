#!/bin/bash
for v in "$@"; do
  lib=""; [ "$v" != "default" ] && lib="$PWD/heat_b200/variants/libhk_$v.so"
  HK_LIB=$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload config5 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('c5 $v', round(j['ms_per_step'],3), round(j['roofline']['kernel_ms_avg'],3), round(j['roofline']['frac'],3), j['roofline']['kernel'], j['parity']['ok'])
"
done
