for i in 1 2 3; do python -m pytest tests -m gpu -x -q 2>&1 | tail -2; done
for i in 1 2; do HB_GEMM_IMPL=fp32 python -m pytest tests -m gpu -x -q 2>&1 | tail -2; done
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01j.json 2> gpurun_out/bench_r01j.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r01j.json')); print(round(d['value']), round(d['ms_per_step'],2), d['config']['phases_ms'], 'e2e', round(d['e2e']['value']), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
( time python bench.py --impl reference --steps 1 --warmup 1 ) 2>&1 | tail -5 | cut -c1-400
