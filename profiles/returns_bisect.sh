#!/bin/bash
# Which part of the stack, if any, changes the learning curve on simple_spread: seeds, fused tcgen05 kernels vs layer-wise,
# pure FP32 SIMT, zero-copy CUDA-graph rollout vs the generic loop.  3 M env steps each (~10 s).
set -x
S=${STEPS:-3000000}
run() { name=$1; shift; env "$@" timeout 900 python examples/returns_mpe.py --impl ours --steps $S --log-interval 5 $EXTRA --out gpurun_out/returns_bisect_$name.json > /dev/null 2> gpurun_out/returns_bisect_$name.err; python - <<PY
import json
d=json.load(open("gpurun_out/returns_bisect_$name.json")); c=d["train_episode_rewards"]
print("$name", "first", round(d["first_10pct_mean"],2), "last", round(d["last_10pct_mean"],2), "mid", round(c[len(c)//2][1],2), "steps/s", int(d["env_steps_per_s"]))
PY
}
EXTRA="--seed 1" run seed1 A=1
EXTRA="--seed 2" run seed2 A=1
EXTRA="--seed 3" run seed3 A=1
EXTRA="--seed 1" run nofused HB_FUSED=0
EXTRA="--seed 1" run simt HB_GEMM_IMPL=fp32
EXTRA="--seed 1 --generic-rollout" run generic A=1
EXTRA="--seed 1 --generic-rollout" run generic_simt HB_GEMM_IMPL=fp32
