#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_iteration.py -q -p no:cacheprovider -x 2>&1 | tail -8 > gpurun_out/r16_gpu.log; tail -6 gpurun_out/r16_gpu.log
timeout 200 python profiles/fused_phase_clock.py 2>&1 | tail -16
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/events_c2_r02d.txt > gpurun_out/bench_c2_r02d.json 2> gpurun_out/bench_c2_r02d.err
python -c "import json; d=json.load(open('gpurun_out/bench_c2_r02d.json')); print('C2', d['value'], d['ms_per_step'], d['config']['phases_ms'], 'e2e', d['e2e']['value'])" || tail -5 gpurun_out/bench_c2_r02d.err
head -6 gpurun_out/events_c2_r02d.txt
