set -x
# --- ncu: launch list of a bench step + full captures of the fused update kernel
NCU_ROWS=819200 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_update_kernel -s 1 -c 1 -o gpurun_out/ncu_fused_update_r02 python profiles/ncu_target.py 2 > gpurun_out/ncu_fused2.log 2>&1; tail -3 gpurun_out/ncu_fused2.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_c2_r02.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err; wc -l gpurun_out/launches_c2_r02.csv
# --- the headline line (with e2e and the in-bench cpu baseline = unmodified reference on a 256-thread sample)
timeout 900 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/events_c2_r02.txt > gpurun_out/bench_c2_r02.json 2> gpurun_out/bench_c2_r02.err
python -c "import json; d=json.load(open('gpurun_out/bench_c2_r02.json')); print('C2', d['value'], d['ms_per_step'], d['config']['phases_ms'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['roofline']['kernel'], d['roofline'].get('frac'))"
for W in C1 C1M C3 C5 C4 C2T; do
  timeout 900 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/events_${W}_r02b.txt > gpurun_out/bench_${W}_r02b.json 2> gpurun_out/bench_${W}_r02b.err
  python -c "import json; d=json.load(open('gpurun_out/bench_${W}_r02b.json')); print('$W', d['value'], d['ms_per_step'], d['config']['phases_ms'], 'e2e', d.get('e2e',{}).get('value'))"
done
timeout 600 python examples/returns_mpe.py --impl ours --steps 3000000 --log-interval 5 --out gpurun_out/returns_mpe_ours_r02.json > gpurun_out/returns_ours.log 2>&1; tail -c 400 gpurun_out/returns_ours.log
