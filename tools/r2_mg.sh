#!/bin/bash
# multi-GPU checks: 2-rank parity tests + config-3 bench at N ranks (peer-memory exchange inside the finish kernel)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -q -x 2>&1 | tail -15 > gpurun_out/r2_mr_pytest.log; tail -6 gpurun_out/r2_mr_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-e2e > gpurun_out/r2_mg_n$N.log 2>&1
HK_NO_PEER=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --no-e2e > gpurun_out/r2_mg_n${N}_nccl.log 2>&1
for f in r2_mg_n$N r2_mg_n${N}_nccl; do python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/$f.log") if x.startswith("{")][-1]; j=json.loads(l)
    print("$f", round(j["ms_per_step"],4), round(j["roofline"]["kernel_ms_avg"],4), j["config"]["comm"], j["gpu_launches"], j.get("parity"), (j.get("graph_replay") or {}))
except Exception as e:
    print("$f", "FAILED", e); print(open("gpurun_out/$f.log").read()[-2500:])
PY
done
