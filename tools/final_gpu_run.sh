# Round-end validation on one B200: GPU tests, smoke, the bench lines and the ncu evidence copied into profiles/.
#   gpurun --timeout 2400 -- 'bash tools/final_gpu_run.sh'
set -x
mkdir -p gpurun_out/final
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/final/bench_default_n1.json 2> gpurun_out/final/bench_default_n1.err; tail -c 700 gpurun_out/final/bench_default_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final/bench_reference_n1.json 2>/dev/null; tail -c 300 gpurun_out/final/bench_reference_n1.json
timeout 200 python bench.py --workload config2 --no-cpu > gpurun_out/final/bench_config2.json 2>/dev/null; tail -c 300 gpurun_out/final/bench_config2.json
timeout 200 python bench.py --workload config5 --no-cpu > gpurun_out/final/bench_config5.json 2>/dev/null; tail -c 300 gpurun_out/final/bench_config5.json
timeout 300 python bench.py --workload config4 --no-cpu --steps 5 --warmup 3 > gpurun_out/final/bench_config4.json 2> gpurun_out/final/bench_config4.err; tail -c 900 gpurun_out/final/bench_config4.json
# launch list of the Lloyd step (own kernels only) and one full capture of the dominant kernel
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:lloyd_|reduce_tc|reduce_row|finalize_' -c 40 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/final/launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:lloyd_tc --launch-skip 4 --launch-count 1 -o gpurun_out/final/prof_c3_tc_final -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/final/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --kernel-name regex:cdist_tc --launch-skip 2 --launch-count 1 -o gpurun_out/final/prof_c2_cdist_final -f python bench.py --workload config2 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/final/ncu_cd.log 2>&1
ls -la gpurun_out/final
