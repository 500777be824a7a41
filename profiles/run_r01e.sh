python -m pytest tests -m gpu -x -q 2>&1 | tail -8
HB_GEMM_IMPL=fp32 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --profile-out gpurun_out/profile_c2_r01e.txt > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r01e.json')); print(round(d['value']), round(d['ms_per_step'],2), d['config']['phases_ms']); print(json.dumps(d['roofline']['named_kernels']))"
tail -3 gpurun_out/bench_r01e.err
head -16 gpurun_out/profile_c2_r01e.txt
