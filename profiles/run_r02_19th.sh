#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
B="--steps 8 --warmup 3 --no-cpu-baseline --no-e2e"
set -x
timeout 400 $TR --master-port 29512 tests/dist_check_replicas.py > gpurun_out/dist_check_replicas_${N}gpu_r02.log 2>&1; tail -3 gpurun_out/dist_check_replicas_${N}gpu_r02.log
timeout 400 $TR --master-port 29513 tests/dist_check_global_batch.py > gpurun_out/dist_check_global_batch_${N}gpu_r02.log 2>&1; tail -4 gpurun_out/dist_check_global_batch_${N}gpu_r02.log
timeout 400 $TR --master-port 29514 bench.py --gpus $N $B > gpurun_out/bench_c2_r02_${N}gpu.json 2> gpurun_out/bench_c2_r02_${N}gpu.err
HB_CRITIC_INTERLEAVE=0 timeout 400 $TR --master-port 29515 bench.py --gpus $N $B > gpurun_out/bench_c2_r02_${N}gpu_freestreams.json 2> gpurun_out/bench_c2_r02_${N}gpu_freestreams.err
HB_CRITIC_INTERLEAVE=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c2_r02_1gpu_interleave.json 2> gpurun_out/bench_c2_r02_1gpu_interleave.err
python - <<PY
import json
for f in ("bench_c2_r02_${N}gpu", "bench_c2_r02_${N}gpu_freestreams", "bench_c2_r02_1gpu_interleave"):
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read().strip().splitlines()[-1]); print(f, round(d["value"]), round(d["ms_per_step"], 2), d["scaling"], d["config"].get("phases_ms"))
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/" + f + ".err").read()[-800:])
PY
