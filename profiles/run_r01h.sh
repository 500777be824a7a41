python tests/debug_tc.py 2>&1 | grep -A9 "c3 impl=1" | head -10
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --profile-out gpurun_out/profile_c2_r01h.txt > gpurun_out/bench_r01h.json 2> gpurun_out/bench_r01h.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r01h.json')); print(round(d['value']), round(d['ms_per_step'],2), d['config']['phases_ms'])"
head -8 gpurun_out/profile_c2_r01h.txt
