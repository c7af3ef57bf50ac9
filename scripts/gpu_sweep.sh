mkdir -p gpurun_out
: > gpurun_out/sweep.jsonl
for cfg in "64 512" "64 2048" "64 8192" "128 512" "128 2048" "128 8192" "256 512"; do
  set -- $cfg
  timeout 150 python bench.py --workload attn --steps 3 --warmup 3 --batch 512 --lmax $2 --attn-dim $1 --attn-heads 4 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/sweep.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    try: d=json.loads(l)
    except Exception: print('bad line', l[:100]); continue
    r=d['roofline']; s=d['config']['attn_shape']
    print(s, 'fwd ms %.3f TF %.0f GB/s %.0f | bwd ms %.3f TF %.0f GB/s %.0f | seq/s %.0f' % (r['fwd']['ms_per_launch'], r['fwd']['achieved'], r['fwd']['hbm_gbs_algorithmic'], r['ms_per_launch'], r['achieved'], r['bwd_hbm_gbs_algorithmic'], d['value']))
PY
