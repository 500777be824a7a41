set -x
bash profiles/run_probe_umma.sh > gpurun_out/probe_umma2.log 2>&1; cat gpurun_out/probe_umma2.log
for W in C2T C4; do
  HB_TRPO_JVP_IMPL=1 HB_RNN_IMPL=persistent python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --profile-out gpurun_out/events_${W}_r02_exp.txt > gpurun_out/bench_${W}_r02_exp.json 2> gpurun_out/bench_${W}_exp.err
  python -c "import json; d=json.load(open('gpurun_out/bench_${W}_r02_exp.json')); print('$W exp', d['value'], d['ms_per_step'], d['config']['phases_ms'])"
done
HB_RNN_IMPL=persistent python bench.py --workload C4R --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_C4R_r02_exp.json 2> gpurun_out/bench_C4R_exp.err
python -c "import json; d=json.load(open('gpurun_out/bench_C4R_r02_exp.json')); print('C4R exp', d['value'], d['ms_per_step'], d['config']['phases_ms'])"
python bench.py --impl reference --ref-cuda --steps 3 --warmup 1 > gpurun_out/bench_c2_r02_reference_cuda.json 2> gpurun_out/ref_cuda.err; tail -c 1500 gpurun_out/bench_c2_r02_reference_cuda.json
python bench.py --impl reference --steps 3 --warmup 1 --ref-cross-check > gpurun_out/bench_c2_r02_reference.json 2> gpurun_out/ref.err; tail -c 2500 gpurun_out/bench_c2_r02_reference.json
