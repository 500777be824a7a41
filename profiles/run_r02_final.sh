#!/bin/bash
# Round 2, final 1-GPU batch on the final code: PCIe probe, phase clock, ncu of the fused kernel + launch list, headline and
# every workload, reference arm (CPU) at the bench config, returns sanity run.
mkdir -p gpurun_out
set -x
timeout 120 python profiles/pcie_copy_probe.py > gpurun_out/pcie_copy_probe_r02.txt 2>&1; cat gpurun_out/pcie_copy_probe_r02.txt
timeout 200 python profiles/fused_phase_clock.py > gpurun_out/fused_phase_clock_r02.txt 2>&1; cat gpurun_out/fused_phase_clock_r02.txt
NCU_ROWS=819200 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_update_kernel -s 1 -c 1 -o gpurun_out/ncu_fused_update_r02_final python profiles/ncu_target.py 2 > gpurun_out/ncu_fused3.log 2>&1; tail -2 gpurun_out/ncu_fused3.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_c2_r02.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err; wc -l gpurun_out/launches_c2_r02.csv
timeout 900 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/events_c2_r02.txt > gpurun_out/bench_c2_r02.json 2> gpurun_out/bench_c2_r02.err
python -c "import json; d=json.load(open('gpurun_out/bench_c2_r02.json')); print('C2', d['value'], d['ms_per_step'], d['config']['phases_ms'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])"
for W in C1 C1M C3 C5 C4 C4R C2T; do
  timeout 900 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/events_${W}_r02.txt > gpurun_out/bench_${W}_r02.json 2> gpurun_out/bench_${W}_r02.err
  python -c "import json; d=json.load(open('gpurun_out/bench_${W}_r02.json')); print('$W', d['value'], d['ms_per_step'], d['config']['phases_ms'], 'e2e', d.get('e2e',{}).get('value'))" || tail -3 gpurun_out/bench_${W}_r02.err
done
timeout 600 python examples/returns_mpe.py --impl ours --steps 2000000 --seed 4 --log-interval 5 --out gpurun_out/returns_mpe_ours_seed4_r02.json > /dev/null 2>&1; python -c "import json; d=json.load(open('gpurun_out/returns_mpe_ours_seed4_r02.json')); print('returns seed4', d['first_10pct_mean'], d['last_10pct_mean'], d['env_steps_per_s'])"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c2_r02_reference.json 2> gpurun_out/bench_c2_r02_reference.err; tail -c 400 gpurun_out/bench_c2_r02_reference.json
