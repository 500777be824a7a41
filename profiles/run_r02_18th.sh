#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
B="--steps 8 --warmup 3 --no-cpu-baseline --no-e2e"
set -x
timeout 400 $TR --master-port 29514 bench.py --gpus $N $B > gpurun_out/bench_c2_r02_${N}gpu.json 2> gpurun_out/bench_c2_r02_${N}gpu.err
HB_OVERLAP_CRITIC=0 timeout 400 $TR --master-port 29515 bench.py --gpus $N $B > gpurun_out/bench_c2_r02_${N}gpu_onestream.json 2> gpurun_out/bench_c2_r02_${N}gpu_onestream.err
python - <<PY
import json
for f in ("bench_c2_r02_${N}gpu", "bench_c2_r02_${N}gpu_onestream"):
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read().strip().splitlines()[-1]); print(f, round(d["value"]), round(d["ms_per_step"], 2), d["scaling"], d["config"].get("phases_ms"), d["config"].get("exchanges"))
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/" + f + ".err").read()[-600:])
PY
