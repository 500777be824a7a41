set -x
timeout 300 python -m pytest tests/test_gpu_fused.py -x -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r3_fused.log 2>&1; tail -40 gpurun_out/r3_fused.log
timeout 120 python -m pytest tests/test_mpe_spread.py -x -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r3_mpe.log 2>&1; tail -15 gpurun_out/r3_mpe.log
