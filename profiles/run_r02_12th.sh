#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -30 > gpurun_out/r12_gpu.log; tail -8 gpurun_out/r12_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/events_c2_r02c.txt > gpurun_out/bench_c2_r02c.json 2> gpurun_out/bench_c2_r02c.err
python -c "import json; d=json.load(open('gpurun_out/bench_c2_r02c.json')); print('C2', d['value'], d['ms_per_step'], d['config']['phases_ms'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value']); print({k:(v['avg_us'], v.get('frac'), v.get('frac_of_same_size_copy')) for k,v in d['roofline']['named_kernels'].items()})" || tail -5 gpurun_out/bench_c2_r02c.err
cat gpurun_out/events_c2_r02c.txt
