python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 32768 65536 131072 262144 819200; do
  HB_CHUNK_ROWS=$c python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --profile-out gpurun_out/profile_chunk_$c.txt 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk',$c, round(d['value']), round(d['ms_per_step'],2), d['config']['phases_ms'])"
done
