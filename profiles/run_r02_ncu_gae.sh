#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gae_seg_kernel -s 8 -c 1 -o gpurun_out/ncu_gae_seg_r02 python profiles/gae_target.py > gpurun_out/ncu_gae1.log 2>&1; tail -1 gpurun_out/ncu_gae1.log
HB_GAE_IMPL=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gae_seg_kernel -s 8 -c 1 -o gpurun_out/ncu_gae_scan_r02 python profiles/gae_target.py > gpurun_out/ncu_gae2.log 2>&1; tail -1 gpurun_out/ncu_gae2.log
