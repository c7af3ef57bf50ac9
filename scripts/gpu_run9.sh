mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_umma.py tests/test_gpu_attention.py -m gpu -q --tb=short -p no:cacheprovider -x -k "128" 2>&1 | tail -12 > gpurun_out/pytest9.log
timeout 120 python bench.py --workload attn --steps 5 --warmup 3 --batch 32 --lmax 4096 --attn-dim 128 --no-cpu-baseline > gpurun_out/bench_attn128.log 2>&1
for f in pytest9 bench_attn128; do echo "== $f"; tail -n 4 gpurun_out/$f.log | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    try:
        d=json.loads(l); r=d['roofline']; print(d['config']['workload'][:60], 'ms/step',round(d['ms_per_step'],3),'fwd TF',round(r['fwd']['achieved'],1),'ms',round(r['fwd']['ms_per_launch'],3),'bwd TF',round(r['achieved'],1),'ms',round(r['ms_per_launch'],3), d.get('kernel_ms_per_call'))
    except Exception as e: print(l[:300])
"; done
