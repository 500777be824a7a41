# Round-1 measurement artefacts (run on the B200 box through gpurun from the repo root)
set -x
python bench.py --steps 5 --warmup 3 --profile-out gpurun_out/events_update_phase_c2_r01.txt > gpurun_out/bench_c2_r01.json 2> gpurun_out/bench_c2_r01.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c2_r01_reference.json 2> gpurun_out/bench_c2_r01_reference.err
# launch list of one update phase (second iteration: skip the first iteration's ~330 update-phase launches)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_|rows_|feat_norm|dw_reduce|featnorm|adam|prepare|pack_umma|moments|normalize|valuenorm|head_kernel" -s 340 -c 330 --csv --log-file gpurun_out/launches_update_r01.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/ncu_launches.err
# launch list of rollout steps (graph replay of the second iteration)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_infer|insert_masks|multi_tensor|counter_add|gae_" -s 700 -c 300 --csv --log-file gpurun_out/launches_rollout_r01.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>> gpurun_out/ncu_launches.err
# full captures of the hot kernels at the C2 row count
# (reports are exported as raw CSV and deleted on the box: gpurun copies back at most 64 MiB)
NCU_ROWS=819200 ncu --set full --clock-control none -k regex:"tc_linear|tc_dx|tc_dw|rows_kernel|rows_grad|feat_norm" -c 12 -o /tmp/prof_hot_r01 python profiles/ncu_target.py 1 > gpurun_out/ncu_hot.log 2>&1
ncu -i /tmp/prof_hot_r01.ncu-rep --page raw --csv > gpurun_out/ncu_hot_r01_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:"fused_infer|gae_tiled" -s 10 -c 2 -o /tmp/prof_rollout_r01 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_rollout.log 2>&1
ncu -i /tmp/prof_rollout_r01.ncu-rep --page raw --csv > gpurun_out/ncu_rollout_r01_raw.csv 2>/dev/null
tail -n 2 gpurun_out/ncu_hot.log; tail -n 2 gpurun_out/ncu_rollout.log; du -sh gpurun_out
python -c "
import json; d=json.loads(open('gpurun_out/bench_c2_r01.json').read().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],2), d['config']['phases_ms'], 'e2e', round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value']), d['gpu_launches'], d['clocks'])"
