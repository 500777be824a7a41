#!/bin/bash
# where do the multi-GPU milliseconds go: per-kernel event profile of the update phase at 2 ranks (labels include hb_allreduce_bucket)
mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
set -x
HB_P2P_ALLREDUCE=1 timeout 400 $TR --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --profile-out gpurun_out/events_c2_r02_2gpu.txt > gpurun_out/bench_c2_r02_2gpu_b.json 2> gpurun_out/bench_c2_r02_2gpu_b.err
cat gpurun_out/events_c2_r02_2gpu.txt
python -c "
import json; d=json.loads(open('gpurun_out/bench_c2_r02_2gpu_b.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['phases_ms'], d['config'].get('exchanges'))"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fused.py -q -p no:cacheprovider -k "discrete_actor_grad or value_grad or log_probs" -x 2>&1 | tail -8 > gpurun_out/sanitizer_memcheck_fused_r02.log; tail -5 gpurun_out/sanitizer_memcheck_fused_r02.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -p no:cacheprovider -k "gae or copy_segments" -x 2>&1 | tail -8 > gpurun_out/sanitizer_memcheck_gae_r02.log; tail -5 gpurun_out/sanitizer_memcheck_gae_r02.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -p no:cacheprovider -k "gae_large_vs_oracle and 4096-200 or parallel_scan" -x 2>&1 | tail -8 > gpurun_out/sanitizer_racecheck_gae_r02.log; tail -5 gpurun_out/sanitizer_racecheck_gae_r02.log
