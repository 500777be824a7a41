python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gae or return" 2>&1 | tail -2
for cw in 16 32; do
HB_GAE_CW=$cw python - <<'PY'
import os, json, torch, bench
for C, sets, reps in ((4096, 8, 5), (65536, 2, 3)):
    r = bench.gae_microbench(torch, 200, C, bench.load_peaks(), sets=sets, reps=reps)
    print("CW", os.environ["HB_GAE_CW"], "C", C, {k: round(v, 3) for k, v in r.items() if k in ("avg_us", "achieved", "frac", "same_bytes_copy_us")})
PY
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
