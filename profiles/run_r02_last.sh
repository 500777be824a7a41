#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/r_last_gpu.log; tail -3 gpurun_out/r_last_gpu.log
python -c "
import __graft_entry__ as g
g.smoke(); print('smoke ok')" 2>&1 | tail -1
