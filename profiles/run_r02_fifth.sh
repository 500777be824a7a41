set -x
timeout 1200 python -m pytest tests -q -m gpu --tb=line -p no:cacheprovider > gpurun_out/r5_gpu.log 2>&1; tail -40 gpurun_out/r5_gpu.log
NCU_ROWS=819200 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_update_kernel -s 1 -c 1 -o gpurun_out/ncu_fused_r02 python profiles/ncu_target.py 2 > gpurun_out/ncu_fused.log 2>&1; tail -5 gpurun_out/ncu_fused.log
ls -la gpurun_out/*.ncu-rep
