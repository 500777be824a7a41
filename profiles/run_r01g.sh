python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tests/debug_tc.py 2>&1 | grep -A16 "c3 impl=1" | head -12
ncu --set full --clock-control none --import-source on -k regex:"fused_infer" -s 20 -c 1 -o gpurun_out/prof_infer_r01 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_infer.log 2>&1
tail -2 gpurun_out/ncu_infer.log
