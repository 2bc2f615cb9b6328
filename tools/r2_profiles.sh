#!/bin/bash
# round-2 evidence: ncu full capture + launch list of the headline step, benches of every workload
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:lloyd_tc_kernel --launch-skip 4 -c 1 -o gpurun_out/r02_c3_tc -f python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r02_c3_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e > /dev/null 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:lloyd_dmma_kernel --launch-skip 4 -c 1 -o gpurun_out/r02_c5_dmma -f python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --workload config5 > gpurun_out/r02_c5_ncu.log 2>&1
for w in config3 config5 config2 config4; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $w --ref-gpu-rows 0 > gpurun_out/r02_bench_$w.log 2>&1
  grep "^{" gpurun_out/r02_bench_$w.log | tail -1 > gpurun_out/r02_bench_$w.json
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r02_bench_$w.json"))
    print("$w", round(j["value"],2), round(j["ms_per_step"],3), round(j["roofline"]["kernel_ms_avg"],3), round(j["roofline"]["frac"],3), j["roofline"]["kernel"], (j.get("e2e") or {}).get("value"), (j.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("$w FAILED", e); print(open("gpurun_out/r02_bench_$w.log").read()[-1500:])
PY
done
for d in randn uncentred overlap; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --data $d > gpurun_out/r02_bench_config3_$d.log 2>&1
  grep "^{" gpurun_out/r02_bench_config3_$d.log | tail -1 > gpurun_out/r02_bench_config3_$d.json
done
ls -la gpurun_out/r02_*
