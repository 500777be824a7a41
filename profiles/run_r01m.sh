python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gae or return" 2>&1 | tail -2
for cw in 8 16 32; do
HB_GAE_CW=$cw python - <<'PY'
import os, json, torch, bench
print("CW", os.environ["HB_GAE_CW"], {k: round(v, 3) if isinstance(v, float) else v for k, v in bench.gae_microbench(torch, 200, 4096, bench.load_peaks()).items() if k in ("avg_us", "achieved", "frac", "same_bytes_copy_us")})
PY
done
