#!/bin/bash
mkdir -p gpurun_out
(cd _r1 && timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > ../gpurun_out/r2c_bench_r1code.log 2>&1)
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c_bench_new.log 2>&1
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum -k regex:lloyd_tc_kernel -c 3 --clock-control none --csv --log-file gpurun_out/r2c_ncu_new.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1
(cd _r1 && timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum -k regex:lloyd_tc_kernel -c 3 --clock-control none --csv --log-file ../gpurun_out/r2c_ncu_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1)
for f in r1code new; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2c_bench_$f.log") if x.startswith("{")][-1]; j=json.loads(l)
    print("$f", round(j["ms_per_step"],3), round(j["roofline"]["kernel_ms_avg"],3), round(j["roofline"]["frac"],3), j["roofline"]["kernel"])
except Exception as e:
    print("$f", "FAILED", e); print(open("gpurun_out/r2c_bench_$f.log").read()[-1500:])
PY
done
grep -v "^==" gpurun_out/r2c_ncu_new.csv | tail -13; grep -v "^==" gpurun_out/r2c_ncu_r1.csv | tail -13
