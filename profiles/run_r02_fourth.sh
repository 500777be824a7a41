set -x
timeout 900 python -m pytest tests -x -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r4_gpu.log 2>&1; tail -25 gpurun_out/r4_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/events_c2_r02_fused.txt > gpurun_out/bench_c2_r02_fused.json 2> gpurun_out/bench_c2_fused.err; tail -5 gpurun_out/bench_c2_fused.err
python -c "import json; d=json.load(open('gpurun_out/bench_c2_r02_fused.json')); print('C2', d['value'], d['ms_per_step'], d['config']['phases_ms'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['avg_us'])"
head -30 gpurun_out/events_c2_r02_fused.txt
