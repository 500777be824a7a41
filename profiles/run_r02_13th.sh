#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/r13_gpu.log; tail -12 gpurun_out/r13_gpu.log
