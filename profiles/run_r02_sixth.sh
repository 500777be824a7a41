set -x
timeout 300 python -m pytest tests/test_gpu_fused.py -q -m gpu --tb=line -p no:cacheprovider 2>&1 | tail -6
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --profile-out gpurun_out/events_c2_r02_fused2.txt > gpurun_out/bench_c2_r02_fused2.json 2> gpurun_out/bench_c2_fused2.err
python -c "import json; d=json.load(open('gpurun_out/bench_c2_r02_fused2.json')); print('C2', d['value'], d['ms_per_step'], d['config']['phases_ms'])"
head -8 gpurun_out/events_c2_r02_fused2.txt
