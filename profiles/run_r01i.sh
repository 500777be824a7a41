python -m pytest tests -m gpu -x -q 2>&1 | tail -5
HB_GEMM_IMPL=fp32 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r01i.json 2> gpurun_out/bench_r01i.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r01i.json')); print(round(d['value']), round(d['ms_per_step'],2), d['config']['phases_ms'])"
