#!/bin/bash
# round-2 evidence on one GPU: GPU test-suite, default bench line, ncu full captures of the final kernels, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu.log; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/r02_bench_default.log 2>&1; echo "bench rc=$?"
grep "^{" gpurun_out/r02_bench_default.log | tail -1 > gpurun_out/r02_bench_default_n1.json
python - <<PY
import json
j=json.load(open("gpurun_out/r02_bench_default_n1.json"))
print("default", round(j["value"],1), j["ms_per_step"], j["roofline"]["kernel_ms_avg"], round(j["roofline"]["frac"],3), j["clocks"], j["graph_replay"]["eager_profiled_ms_per_step"])
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:lloyd_tc_kernel --launch-skip 4 -c 1 -o gpurun_out/r02_c3_tc -f python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-extras > gpurun_out/r02_c3_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lloyd|finish|finalize|reduce" -c 40 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-extras > /dev/null 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:lloyd_dmma_kernel --launch-skip 4 -c 1 -o gpurun_out/r02_c5_dmma -f python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --workload config5 > gpurun_out/r02_c5_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:lloyd_tc_kernel --launch-skip 4 -c 1 -o gpurun_out/r02_c3_tc_randn -f python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-extras --data randn > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
