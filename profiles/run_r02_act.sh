#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -6 > gpurun_out/r_act_gpu.log; tail -4 gpurun_out/r_act_gpu.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_r02e.json 2> gpurun_out/bench_c2_r02e.err
python -c "import json; d=json.load(open('gpurun_out/bench_c2_r02e.json')); print('C2', d['value'], d['ms_per_step'], d['config']['phases_ms'], 'e2e', d['e2e']['value'])" || tail -5 gpurun_out/bench_c2_r02e.err
