python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r01n.json 2> gpurun_out/bench_r01n.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r01n.json').read().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],2), d['config']['phases_ms']); g=d['roofline']['named_kernels']; print({k:(round(v['avg_us'],2), round(v['frac'],3)) for k,v in g.items()})"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"fused_infer" -s 30 -c 3 --csv --log-file gpurun_out/infer_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
grep -v "^==" gpurun_out/infer_metrics.csv | cut -d, -f5,13- | tail -9
