./profiles/_probe_mn 2>&1 | tail -6
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/profile_c2_r01d.txt > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r01d.json')); print(round(d['value']), round(d['ms_per_step'],2), d['config']['phases_ms'], d['e2e']['value'], d['e2e']['ms_per_step'])"
tail -3 gpurun_out/bench_r01d.err
head -14 gpurun_out/profile_c2_r01d.txt
NCU_ROWS=409600 ncu --set full --clock-control none --import-source on -k regex:"tc_linear|tc_dx|tc_dw|rows_kernel|rows_grad" -c 8 -o gpurun_out/prof_bigm_r01 python profiles/ncu_target.py 1 > gpurun_out/ncu_bigm.log 2>&1
tail -2 gpurun_out/ncu_bigm.log
