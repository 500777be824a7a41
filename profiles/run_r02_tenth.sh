#!/bin/bash
# Round 2, tenth GPU batch: new parity tests (activation goldens, minibatch / MAPPO train replays, checkpoints, MPE rollout),
# GAE kernel variants, what changes the simple_spread learning curve, C3 / C5 benches after the rollout fix.
mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r10_gpu.log; cat gpurun_out/r10_gpu.log | tail -25
timeout 600 python profiles/gae_variants.py > gpurun_out/gae_variants_r02.txt 2>&1; cat gpurun_out/gae_variants_r02.txt
bash profiles/returns_bisect.sh 2>&1 | grep -v "^+" | tee gpurun_out/returns_bisect_r02.txt
for W in C3 C5; do
  timeout 900 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/events_${W}_r02b.txt > gpurun_out/bench_${W}_r02b.json 2> gpurun_out/bench_${W}_r02b.err
  python -c "import json; d=json.load(open('gpurun_out/bench_${W}_r02b.json')); print('$W', d['value'], d['ms_per_step'], d['config']['phases_ms'], 'e2e', d.get('e2e',{}).get('value'))" || tail -5 gpurun_out/bench_${W}_r02b.err
done
