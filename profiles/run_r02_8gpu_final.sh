#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
set -x
timeout 200 $TR --master-port 29511 tests/dist_check_allreduce.py > gpurun_out/dist_check_allreduce_${N}gpu.log 2>&1; tail -1 gpurun_out/dist_check_allreduce_${N}gpu.log
timeout 300 $TR --master-port 29514 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_r02_${N}gpu.json 2> gpurun_out/bench_c2_r02_${N}gpu.err
python - <<PY
import json
for f in ("bench_c2_r02_${N}gpu",):
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read().strip().splitlines()[-1]); print(f, round(d["value"]), round(d["ms_per_step"], 2), d["scaling"], d["config"].get("phases_ms"), d["config"].get("exchanges"), "e2e", d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/" + f + ".err").read()[-800:])
PY
