python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r01l.json 2> gpurun_out/bench_r01l.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r01l.json').read().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],2), d['config']['phases_ms']); g=d['roofline']['named_kernels']['gae']; print(g['avg_us'], g['achieved'], g['frac'], g['same_bytes_copy_us'])"
