#!/bin/bash
# end-of-round validation on one GPU: full GPU test-suite, smoke(), default bench (with extras)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu.log; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 1200 python bench.py > gpurun_out/r02_bench_default.log 2>&1; echo "bench rc=$?"
grep "^{" gpurun_out/r02_bench_default.log | tail -1 > gpurun_out/r02_bench_default_n1.json
python - <<PY
import json
j=json.load(open("gpurun_out/r02_bench_default_n1.json"))
print("default", round(j["value"],1), j["ms_per_step"], j["roofline"]["kernel_ms_avg"], round(j["roofline"]["frac"],3), j["clocks"], "e2e", j["e2e"]["value"], "cpu", j["cpu_baseline"]["value"], j["cpu_baseline"]["kind"], j.get("reference_gpu"), j.get("parity"))
for k,v in j.get("other_workloads",{}).items(): print(" ", k, v)
PY
