python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/profile_c2_r01k.txt > gpurun_out/bench_r01k.json 2> gpurun_out/bench_r01k.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r01k.json').read().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],2), d['config']['phases_ms'], 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],1), d['gpu_launches'])"
head -12 gpurun_out/profile_c2_r01k.txt
