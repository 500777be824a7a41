#!/bin/bash
# Round 2, N-GPU batch (gpurun --gpus N, default 8): peer-memory allreduce check + latency, weak scaling of C2 with the
# one-shot exchange and with NCCL, C5 at 1024 threads per GPU (BASELINE configs[4]: 8192 threads over 8 GPUs), strong
# scaling of C2 and C4 (the workload's thread count split over the ranks).
mkdir -p gpurun_out
N=${NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
B="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
set -x
nvidia-smi topo -m 2>&1 | head -12 > gpurun_out/topo_${N}gpu.txt
timeout 300 $TR --master-port 29511 tests/dist_check_allreduce.py > gpurun_out/dist_check_allreduce_${N}gpu.log 2>&1; tail -2 gpurun_out/dist_check_allreduce_${N}gpu.log
timeout 400 $TR --master-port 29514 bench.py --gpus $N $B > gpurun_out/bench_c2_r02_${N}gpu.json 2> gpurun_out/bench_c2_r02_${N}gpu.err
HB_P2P_ALLREDUCE=0 timeout 400 $TR --master-port 29515 bench.py --gpus $N $B > gpurun_out/bench_c2_r02_${N}gpu_nccl.json 2> gpurun_out/bench_c2_r02_${N}gpu_nccl.err
timeout 400 $TR --master-port 29516 bench.py --gpus $N --workload C5 $B > gpurun_out/bench_c5_r02_${N}gpu.json 2> gpurun_out/bench_c5_r02_${N}gpu.err
timeout 400 $TR --master-port 29517 bench.py --gpus $N --workload C2 --scaling strong $B > gpurun_out/bench_c2_r02_${N}gpu_strong.json 2> gpurun_out/bench_c2_r02_${N}gpu_strong.err
timeout 400 $TR --master-port 29518 bench.py --gpus $N --workload C4 --scaling strong $B > gpurun_out/bench_c4_r02_${N}gpu_strong.json 2> gpurun_out/bench_c4_r02_${N}gpu_strong.err
python - <<PY
import json
for f in ("bench_c2_r02_${N}gpu", "bench_c2_r02_${N}gpu_nccl", "bench_c5_r02_${N}gpu", "bench_c2_r02_${N}gpu_strong", "bench_c4_r02_${N}gpu_strong"):
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read().strip().splitlines()[-1]); print(f, round(d["value"]), round(d["ms_per_step"], 2), d["scaling"], d["config"].get("phases_ms"), d["config"].get("exchanges"))
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/" + f + ".err").read()[-600:])
PY
