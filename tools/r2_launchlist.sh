#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lloyd_|finish_kernel|finalize_kernel|reduce_tc" -c 24 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-extras > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lloyd_|finish_kernel|finalize_kernel|reduce_tc" -c 24 --csv --log-file gpurun_out/r02_launches_c3_shard8.csv python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-extras --rows 12500000 > /dev/null 2>&1
grep -v "^==" gpurun_out/r02_launches_c3_shard8.csv | python -c "
import csv,sys
for r in csv.reader(sys.stdin):
    if len(r)>10 and r[0].isdigit(): print(r[0], r[4][:70], r[-1])
" | tail -12
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-extras --rows 12500000 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('shard8 on 1 GPU: step', j['ms_per_step'], 'kernel', j['roofline']['kernel_ms_avg'], 'eager', j['graph_replay']['eager_profiled_ms_per_step'])
"
