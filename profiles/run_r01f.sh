python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --profile-out gpurun_out/profile_c2_r01f.txt > gpurun_out/bench_r01f.json 2> gpurun_out/bench_r01f.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r01f.json')); print(round(d['value']), round(d['ms_per_step'],2), d['config']['phases_ms']); print(json.dumps(d['roofline']['named_kernels']['gae']))"
tail -3 gpurun_out/bench_r01f.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_infer|insert_masks|gae_tiled|multi_tensor|counter_add" -c 40 --csv --log-file gpurun_out/launches_rollout.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_rollout.csv')) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
d=collections.defaultdict(list)
for r in rows[1:]:
    d[r[ki][:60]].append(float(r[vi].replace(',','')))
for k,v in d.items(): print(k, len(v), sum(v)/len(v))
PY
