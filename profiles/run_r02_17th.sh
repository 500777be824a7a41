#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/r17_gpu.log; tail -4 gpurun_out/r17_gpu.log
NCU_ROWS=819200 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_update_kernel -s 1 -c 1 -o gpurun_out/ncu_fused_update_r02_final python profiles/ncu_target.py 2 > gpurun_out/ncu_fused3.log 2>&1; tail -2 gpurun_out/ncu_fused3.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_c2_r02.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err; wc -l gpurun_out/launches_c2_r02.csv
timeout 200 python profiles/fused_phase_clock.py > gpurun_out/fused_phase_clock_r02.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/events_c2_r02.txt > gpurun_out/bench_c2_r02.json 2> gpurun_out/bench_c2_r02.err
python -c "import json; d=json.load(open('gpurun_out/bench_c2_r02.json')); print('C2', d['value'], d['ms_per_step'], d['config']['phases_ms'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])"
for W in C1 C1M C2T; do
  timeout 900 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/events_${W}_r02.txt > gpurun_out/bench_${W}_r02.json 2> gpurun_out/bench_${W}_r02.err
  python -c "import json; d=json.load(open('gpurun_out/bench_${W}_r02.json')); print('$W', d['value'], d['ms_per_step'], d['config']['phases_ms'], 'e2e', d.get('e2e',{}).get('value'))" || tail -3 gpurun_out/bench_${W}_r02.err
done
python -c "
import __graft_entry__ as g
g.smoke(); print('smoke ok')" 2>&1 | tail -2
