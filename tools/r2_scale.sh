#!/bin/bash
# scaling point: config-3 bench at N ranks
N=${1:-8}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r2_scale_n$N.log 2>&1
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-e2e > gpurun_out/r2_scale_n$N.log 2>&1
fi
python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2_scale_n$N.log") if x.startswith("{")][-1]; j=json.loads(l)
    print("N=$N", "value", round(j["value"],1), "ms/step", round(j["ms_per_step"],4), "kernel", round(j["roofline"]["kernel_ms_avg"],4), "frac", round(j["roofline"]["frac"],3), j["config"]["comm"], "launches", j["gpu_launches"], j.get("parity"), j.get("graph_replay"))
except Exception as e:
    print("N=$N", "FAILED", e); print(open("gpurun_out/r2_scale_n$N.log").read()[-2500:])
PY
