#!/bin/bash
# Round 2, multi-GPU batch (gpurun --gpus N): one-shot peer-memory allreduce checks + latency, replica / global-batch checks,
# weak-scaling bench at N ranks with the peer-memory exchange and with NCCL.
mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
set -x
timeout 300 python -m pytest tests/test_gpu_mpe_rollout.py tests/test_gpu_rnn.py tests/test_gpu_kernels.py -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/r11_gpu.log; tail -5 gpurun_out/r11_gpu.log
nvidia-smi topo -m 2>&1 | head -14 > gpurun_out/topo_${N}gpu.txt
timeout 300 $TR --master-port 29511 tests/dist_check_allreduce.py > gpurun_out/dist_check_allreduce_${N}gpu.log 2>&1; tail -4 gpurun_out/dist_check_allreduce_${N}gpu.log
timeout 600 $TR --master-port 29512 tests/dist_check_replicas.py > gpurun_out/dist_check_replicas_${N}gpu_r02.log 2>&1; tail -4 gpurun_out/dist_check_replicas_${N}gpu_r02.log
timeout 600 $TR --master-port 29513 tests/dist_check_global_batch.py > gpurun_out/dist_check_global_batch_${N}gpu_r02.log 2>&1; tail -6 gpurun_out/dist_check_global_batch_${N}gpu_r02.log
timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_r02_${N}gpu.json 2> gpurun_out/bench_c2_r02_${N}gpu.err; tail -c 600 gpurun_out/bench_c2_r02_${N}gpu.json
HB_P2P_ALLREDUCE=0 timeout 600 $TR --master-port 29515 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_r02_${N}gpu_nccl.json 2> gpurun_out/bench_c2_r02_${N}gpu_nccl.err; tail -c 600 gpurun_out/bench_c2_r02_${N}gpu_nccl.json
python - <<PY
import json
for f in ("gpurun_out/bench_c2_r02_${N}gpu.json", "gpurun_out/bench_c2_r02_${N}gpu_nccl.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["config"].get("phases_ms"), d["config"].get("exchanges"))
    except Exception as e:
        print(f, "FAILED", e)
PY
