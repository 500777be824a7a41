set -x
timeout 900 python -m pytest tests -q -m gpu --tb=line -p no:cacheprovider > gpurun_out/r8_gpu.log 2>&1; tail -12 gpurun_out/r8_gpu.log
timeout 120 python profiles/fused_phase_clock.py 2>&1 | tail -18
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --profile-out gpurun_out/events_c2_r02_fused4.txt > gpurun_out/bench_c2_r02_fused4.json 2> gpurun_out/bench_c2_fused4.err
python -c "import json; d=json.load(open('gpurun_out/bench_c2_r02_fused4.json')); print('C2', d['value'], d['ms_per_step'], d['config']['phases_ms'])"
head -8 gpurun_out/events_c2_r02_fused4.txt
