# First GPU minutes of round 2 (run through gpurun from the repo root): what round 1 could not run any more.
set -x
mkdir -p gpurun_out
# 1. the Discrete(12) head dispatch fix + the synthetic-SMAC-width goldens
python -m pytest tests/test_gpu_zz_wide_heads.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r2_wide.log 2>&1; tail -5 gpurun_out/r2_wide.log
# 2. the two kernels written blind (default off): persistent GRU recurrence, tcgen05 tangent block
HB_RUN_EXPERIMENTAL=1 python -m pytest tests/test_gpu_rnn.py tests/test_gpu_zz_wide_heads.py -q -m gpu --tb=short -p no:cacheprovider \
    -k "persistent or tensor_core" > gpurun_out/r2_experimental.log 2>&1; tail -15 gpurun_out/r2_experimental.log
# 3. re-measure the secondary workloads (C4 was last timed with the dispatch bug present)
for W in C2T C4R C4; do
  python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --profile-out gpurun_out/events_${W}_r02.txt \
      > gpurun_out/bench_${W}_r02.json 2> gpurun_out/bench_${W}.err
  python -c "import json; d=json.load(open('gpurun_out/bench_${W}_r02.json')); print('$W', d['value'], d['ms_per_step'], d['config']['phases_ms'])"
done
