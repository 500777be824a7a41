set -x
timeout 300 python -m pytest tests/test_gpu_fused.py -q -m gpu --tb=line -p no:cacheprovider 2>&1 | tail -6
timeout 120 python profiles/fused_phase_clock.py 2>&1 | tail -18
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --profile-out gpurun_out/events_c2_r02_fused3.txt > gpurun_out/bench_c2_r02_fused3.json 2> gpurun_out/bench_c2_fused3.err
python -c "import json; d=json.load(open('gpurun_out/bench_c2_r02_fused3.json')); print('C2', d['value'], d['ms_per_step'], d['config']['phases_ms'])"
head -6 gpurun_out/events_c2_r02_fused3.txt
